"""CPU oracle package -- TEST INFRASTRUCTURE ONLY (see hx_oracle.py header)."""
