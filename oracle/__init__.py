"""CPU oracle for the CLR hot path.  TEST INFRASTRUCTURE ONLY -- see clr_oracle.py."""
