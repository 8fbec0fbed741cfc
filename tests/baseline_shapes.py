"""The BASELINE.json configurations at their own grid shapes: the builders live in chimera_b200/synthetic.py (bench.py
--config uses them too); re-exported here for the tests."""
from chimera_b200.synthetic import (baseline_case, baseline_species, c1a_fel, c1b_lpa, c2_space_charge, c3_lwfa,  # noqa: F401
                                    fill_plasma, gaussian_beam)
