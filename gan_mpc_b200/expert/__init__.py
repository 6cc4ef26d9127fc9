"""Expert proposal network (reference expert/): nn.py (network shells + flat layout),
expert_model.py (ExpertModel over gmpc_expert_propose), synthetic_expert.py (seeded stand-in)."""

from gan_mpc_b200.expert.synthetic_expert import SyntheticExpert  # noqa: F401
