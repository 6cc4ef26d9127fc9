"""CriticModel shell (reference critic/critic_model.py:6-16)."""

from gan_mpc_b200 import base


class CriticModel(base.BaseCriticModel):
    def __init__(self, config, model):
        self.config = config
        self.model = model

    def init(self, *args, device="cuda"):
        return self.model.init(*self.model.get_init_params(*args), device=device)

    def predict(self, xseq, params):
        raise NotImplementedError(
            "CriticModel.predict is evaluated inside libgmpc (gmpc_critic_forward); call it "
            "through JS_MPC.critic_logits / critic_loss")
