"""Interface shells carrying the reference's class names (base.py:4-49 of returaj/gan_mpc), so that
`isinstance` checks and subclassing written against the reference keep working.  They hold no
arithmetic: every abstract entry point raises NotImplementedError naming the class and method."""


def _abstract(method_name):
    def method(self, *args, **kwargs):
        raise NotImplementedError(f"{type(self).__name__}.{method_name} is abstract")
    method.__name__ = method_name
    return method


def _shell(class_name, abstract_methods, parent=object, keeps_config=False):
    namespace = {name: _abstract(name) for name in abstract_methods}
    if keeps_config:
        def __init__(self, config):
            self.config = config
        namespace["__init__"] = __init__
    namespace["__doc__"] = f"abstract shell: {', '.join(abstract_methods)}"
    return type(class_name, (parent,), namespace)


# models own a config and expose init(...) plus one evaluation entry point
BaseCostModel = _shell("BaseCostModel", ("init", "get_cost"), keeps_config=True)
BaseDynamicsModel = _shell("BaseDynamicsModel", ("init", "predict"), keeps_config=True)
BaseCriticModel = _shell("BaseCriticModel", ("init", "predict"), keeps_config=True)
# network shells describe how their parameters are initialised
BaseNN = _shell("BaseNN", ("get_init_params",))
BaseCostNN = _shell("BaseCostNN", ("get_cost",), parent=BaseNN)
BaseDynamicsNN = _shell("BaseDynamicsNN", ("get_carry",), parent=BaseNN)
