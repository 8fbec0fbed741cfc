"""chimera_b200 -- B200-native (sm_100a) implementation of CHIMERA's PIC-cycle hot path.

Public surface
  chimera_b200.fimera        drop-in for the reference's f2py module ``chimera.moduls.fimera``
                             (host numpy buffers in, CUDA kernels underneath, no CPU fallback)
  chimera_b200.solver_setup  host-side builder of the DHT / mode-coupling / PSATD tables
  chimera_b200.engine        device-resident PIC step (particles and fields stay in HBM)
  chimera_b200.build         compiles csrc/ into libchimera_b200.so with nvcc for sm_100a
"""
__version__ = "0.1.0"
