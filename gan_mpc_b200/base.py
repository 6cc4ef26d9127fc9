"""Abstract bases with the reference's class names (base.py:4-49 of returaj/gan_mpc), so that `isinstance`
checks and subclassing written against the reference keep working.  No arithmetic lives here."""


class BaseCostModel:
    def __init__(self, config):
        self.config = config

    def init(self, *args):
        raise NotImplementedError

    def get_cost(self, xc, u, t, params, *args):
        raise NotImplementedError


class BaseDynamicsModel:
    def __init__(self, config):
        self.config = config

    def init(self, *args):
        raise NotImplementedError

    def predict(self, xc, u, t, params, *args):
        raise NotImplementedError


class BaseCriticModel:
    def __init__(self, config):
        self.config = config

    def init(self, *args):
        raise NotImplementedError

    def predict(self, xseq, params):
        raise NotImplementedError


class BaseNN:
    def get_init_params(self, *args):
        raise NotImplementedError


class BaseCostNN(BaseNN):
    def get_cost(self, params, x):
        raise NotImplementedError


class BaseDynamicsNN(BaseNN):
    def get_carry(self, x):
        raise NotImplementedError
