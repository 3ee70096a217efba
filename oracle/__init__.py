"""CPU oracle (test infrastructure only) — see nbhood_oracle.py."""
