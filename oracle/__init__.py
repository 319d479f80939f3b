"""CPU oracle for the BGP hot path -- TEST INFRASTRUCTURE.  Importable only from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
