"""CPU oracle for the gan_mpc planning hot path.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the reference (returaj/gan_mpc) ships no tests, golden vectors
or fixtures, and its JAX stack (jax 0.4.13 / flax 0.7.2 / optax 0.1.7 /
trajax@c94a637) is neither installed nor installable offline, so this
restatement cannot be checked against outputs of the reference itself.  Its only
pins are self-consistency checks (autograd vs hand-written adjoint, finite
differences, fp32 vs fp64) and the frozen vectors under tests/golden/ that it
generated itself.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package.  The product (gan_mpc_b200/) never does.
"""
