"""Test-infrastructure oracle (see bin_oracle.py header). Never imported by bin_b200."""
