"""Oracle = CPU restatements of the reference path.  TEST INFRASTRUCTURE ONLY:
nothing under eigensolver_b200/ imports this package."""
