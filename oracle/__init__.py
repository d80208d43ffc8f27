"""oracle -- CPU restatement of the reference's nonbonded path.  TEST INFRASTRUCTURE ONLY: may be
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
