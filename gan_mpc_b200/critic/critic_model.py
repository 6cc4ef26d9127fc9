"""CriticModel with the reference's name and constructor (critic/critic_model.py:6-16).  The LSTM
discriminator itself is evaluated by libgmpc (csrc/critic.cuh): this object only owns the network
shell (hyper-parameters, flat <-> pytree layout) and builds initial parameters."""

from gan_mpc_b200 import base


class CriticModel(base.BaseCriticModel):
    def __init__(self, config, model):
        super().__init__(config)
        self.model = model

    def init(self, *args, device="cuda"):
        """args = (seed, x_size), as assembled by gan.runner.get_params."""
        init_args = self.model.get_init_params(*args)
        return self.model.init(*init_args, device=device)

    def predict(self, xseq, params):
        """critic/critic_model.py:15-16 is a jitted model.apply; here the score comes from the kernel."""
        raise NotImplementedError(
            "CriticModel.predict is evaluated inside libgmpc (gmpc_critic_forward); call it "
            "through JS_MPC.critic_logits / critic_loss")
