"""CPU oracle for the CHIMERA PIC hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  See chimera_oracle.cpp for the "parity unpinned" statement.
"""
