"""CPU oracle for the flow-state NF-MCMC hot path.  TEST INFRASTRUCTURE ONLY.

Everything under oracle/ is a CPU restatement of the reference's algorithm
(numpy / torch-CPU / plain C), each function citing the reference file:line it
follows.  It is the *checker*: only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import it.  The product
package (flowstate_b200/) never imports oracle/ and has no CPU fallback.

Pinning: the reference holds no golden vectors for this path, so
the oracle is pinned against outputs of the reference itself, generated in the
build container by oracle/make_golden.py (which imports /root/reference
unmodified) and committed under tests/golden/.  tests/test_oracle_golden.py
re-checks the oracle against those fixtures on every run.
"""
