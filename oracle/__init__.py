"""Parity oracle (test infrastructure only).  See oracle/pangu_oracle.py."""
