"""CPU oracle for the SamCon hot path -- test infrastructure only (see samcon_oracle.c)."""
