"""TEST INFRASTRUCTURE ONLY. CPU restatement of the reference algorithm (see siglip_oracle.py).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this."""
