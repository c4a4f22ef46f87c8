"""CPU oracle -- test infrastructure only (see oracle.c)."""
