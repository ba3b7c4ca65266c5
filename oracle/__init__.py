"""CPU oracle package -- test infrastructure only (see kzg_oracle.c / kzg_oracle.py headers)."""
