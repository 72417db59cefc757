"""CPU oracle -- TEST INFRASTRUCTURE.  Importable only from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs (see oracle/nomp_oracle.c)."""
