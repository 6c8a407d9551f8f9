"""Test infrastructure only (see af_oracle.py header).  Never imported by audioflux_b200/."""
