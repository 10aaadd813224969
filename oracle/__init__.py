"""TEST INFRASTRUCTURE ONLY - CPU oracle of the Lighthouse 2 render path (see oracle/*.h headers).
Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs only."""
