"""CPU oracle package -- test infrastructure only (see foam_oracle.cpp header). PARITY UNPINNED."""
