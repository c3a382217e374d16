"""TEST INFRASTRUCTURE ONLY: parity oracle for ilswiss_b200.  Never imported by the product path."""
