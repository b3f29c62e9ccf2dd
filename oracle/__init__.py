"""CPU oracle (test infrastructure only; see cd_oracle.py)."""
