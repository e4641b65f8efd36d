"""oracle/ — TEST INFRASTRUCTURE ONLY (checker for the CUDA path, never the product).

Importable by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs only."""
