"""CPU tests of the host-side mirror of the reference API (no kernels launched)."""

import os

import numpy as np
import pytest
import torch
import yaml

from gan_mpc_b200 import parallel, synthetic, utils
from gan_mpc_b200.config import load_config
from gan_mpc_b200.critic import nn as critic_nn
from gan_mpc_b200.policy import optimizers as opt
from oracle import critic as ocritic

REF = "/root/reference/config"


@pytest.mark.parametrize("name", ["l2_hyperparameters.yaml", "gan_hyperparameters.yaml"])
def test_yaml_matches_reference_values(name):
    """Our hyper-parameter files carry the reference's keys and values (+ the mpc.planner block)."""
    ours = yaml.safe_load(open(os.path.join(load_config.CONFIG_DIR, name)))
    assert ours["mpc"]["planner"]["method"] == "ilqr"   # the reference's planner is the default
    if not os.path.exists(os.path.join(REF, name)):
        pytest.skip("reference tree not mounted")
    ref = yaml.safe_load(open(os.path.join(REF, name)))

    def check(r, o, path=""):
        for k, v in r.items():
            if path + k == "expert_prediction":
                continue                      # expert network: out of scope
            assert k in o, f"missing key {path}{k}"
            if isinstance(v, dict):
                check(v, o[k], path + k + ".")
            else:
                assert o[k] == v, f"{path}{k}: {o[k]} != {v}"
    check(ref, ours)


def test_config_attribute_tree_roundtrip():
    c = load_config.Config.from_dict({"a": 1, "b": {"c": [1, 2], "d": {"e": "x"}}})
    assert c.a == 1 and c.b.c == [1, 2] and c.b.d.e == "x"
    assert c.to_dict() == {"a": 1, "b": {"c": [1, 2], "d": {"e": "x"}}}


def test_critic_flatten_roundtrip_and_layout():
    n, F, L, H = 5, 8, 3, 6
    model = critic_nn.LSTM(F, L, H)
    flat = torch.from_numpy(synthetic.critic_params_flat(0, n, F, L, H))
    assert model.param_count(n) == flat.numel() == ocritic.critic_param_count(n, F, L, H)
    tree = model.unflatten(flat, n)
    cell = tree["params"][critic_nn.CELL]
    assert set(cell) == {"ii", "if", "ig", "io", "hi", "hf", "hg", "ho"}       # flax OptimizedLSTMCell
    assert "bias" not in cell["ii"] and cell["hi"]["bias"].shape == (F,)
    assert cell["ig"]["kernel"].shape == (n, F) and cell["ho"]["kernel"].shape == (F, F)
    assert tree["params"]["Dense_2"]["kernel"].shape == (H, 1)
    assert torch.equal(model.flatten(tree), flat)
    # the oracle reads the same flat layout
    o = ocritic.unflatten(flat, n, F, L, H)
    assert torch.equal(o["Wi"][:, 2 * F:3 * F], cell["ig"]["kernel"])


def test_lecun_normal_statistics():
    rng = np.random.Generator(np.random.PCG64(0))
    W = synthetic.lecun_normal(rng, 400, 300)
    assert abs(W.std() - np.sqrt(1 / 400)) < 0.02 * np.sqrt(1 / 400)          # variance 1/fan_in
    assert np.abs(W).max() <= 2.0 * np.sqrt(1 / 400) / 0.87962566103423978 + 1e-7
    Q = synthetic.orthogonal(rng, 16)
    assert np.allclose(Q.T @ Q, np.eye(16), atol=1e-5)


def test_arbitrary_closures_are_rejected():
    """No CPU fallback: a plain Python cost/dynamics closure cannot run in the planner."""
    with pytest.raises(TypeError, match="no CPU fallback"):
        opt.ilqr_solve(lambda x, u, t, p: 0.0, lambda x, u, t, p: x, None, None, {}, (None,), (), {})
    with pytest.raises(TypeError):
        opt.objective(lambda x, u, t: 0.0, lambda x, u, t: x, None, None)
    f = lambda *a: 0.0
    with pytest.raises(TypeError, match="no CPU fallback"):
        opt.bilevel_optimization(f, f, f, None, None, {}, (None,), (), (None,), {})
    with pytest.raises(TypeError):
        opt.cost_hessian_wrt_control(f, f, None, None)
    with pytest.raises(TypeError, match="no CPU fallback"):
        opt.cost_vjp(f, f, None, None, None, {}, (None,))


def test_model_factories_and_mask_labels():
    c = utils.get_config(os.path.join(load_config.CONFIG_DIR, "gan_hyperparameters.yaml"))
    cost, _ = utils.get_cost_model(c)
    dyn, _ = utils.get_dynamics_model(c, 3)
    critic, _ = utils.get_critic_model(c)
    assert (cost.model.num_layers, cost.model.num_hidden_units, cost.model.fout) == (3, 128, 10)
    assert (dyn.model.num_layers, dyn.model.num_hidden_units, dyn.model.x_out) == (4, 200, 3)
    assert (critic.model.lstm_features, critic.model.num_layers) == (64, 1)
    labels = utils.get_masked_labels(["mpc_weights", "cost_params", "critic_params"],
                                     c.mpc.train.critic.no_grads, "tx", "zero")
    assert labels == {"mpc_weights": "zero", "cost_params": "zero", "critic_params": "tx"}
    p = dyn.init(0, 1, device="cpu")
    assert p["params"]["Dense_0"]["kernel"].shape == (4, 200) and p["params"]["Dense_3"]["kernel"].shape == (200, 3)
    assert float(p["params"]["Dense_1"]["bias"].abs().sum()) == 0.0


def test_timeit_appends_minutes():
    @utils.timeit
    def f():
        return 1, 2
    out = f()
    assert out[:2] == (1, 2) and len(out) == 3 and out[2] >= 0.0


@pytest.mark.parametrize("B,world", [(10, 4), (4096, 8), (3, 8), (0, 2), (262144, 8)])
def test_shard_range_partitions_the_batch(B, world):
    edges = [parallel.shard_range(B, r, world) for r in range(world)]
    assert edges[0][0] == 0 and edges[-1][1] == B
    sizes = [hi - lo for lo, hi in edges]
    assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
    assert max(sizes) - min(sizes) <= 1


def test_params_npy_round_trip(tmp_path):
    """utils.save_all_args / load_params (reference utils.py:135-156): run-id directories,
    config.json, params.npy as a pickled pytree with the flax Dense_i/{kernel,bias} layout."""
    import numpy as np
    import torch
    c = utils.get_config(os.path.join(load_config.CONFIG_DIR, "l2_hyperparameters.yaml"))
    dyn, _ = utils.get_dynamics_model(c, 3)
    params = {"mpc_weights": torch.tensor([-2.0, 3.0, -3.0]), "dynamics_params": dyn.init(0, 1, device="cpu"),
              "expert_params": {}}
    d0 = utils.save_all_args(str(tmp_path / "runs"), params, c.mpc.model.to_dict(), ({"loss": [1.0, 0.5]}, "loss.json"))
    d1 = utils.save_all_args(str(tmp_path / "runs"), params, c.mpc.model.to_dict())
    assert os.path.basename(d0) == "0" and os.path.basename(d1) == "1"
    assert utils.load_json(os.path.join(d0, "loss.json")) == {"loss": [1.0, 0.5]}
    assert utils.load_json(os.path.join(d0, "config.json"))["dynamics"]["mlp"]["num_hidden_units"] == 200
    raw = np.load(os.path.join(d0, "params.npy"), allow_pickle=True).item()
    assert isinstance(raw["dynamics_params"]["params"]["Dense_0"]["kernel"], np.ndarray)
    back = utils.load_params(os.path.join(d0, "params.npy"), device="cpu")
    assert torch.equal(back["mpc_weights"], params["mpc_weights"]) and back["expert_params"] == {}
    for i in range(4):
        for k in ("kernel", "bias"):
            assert torch.equal(back["dynamics_params"]["params"][f"Dense_{i}"][k],
                               params["dynamics_params"]["params"][f"Dense_{i}"][k])


def test_expert_network_layout_and_oracle():
    """flat layout <-> flax-named pytree, parameter counts, and the oracle's call sequence
    (history warm-up then free-running proposal; row 0 of the goals is the current state)."""
    import numpy as np
    import torch
    from gan_mpc_b200 import synthetic
    from gan_mpc_b200.expert import expert_model, nn as expert_nn
    from oracle import expert as oexpert
    c = utils.get_config(os.path.join(load_config.CONFIG_DIR, "l2_hyperparameters.yaml"))
    em = utils.get_expert_model(c, 3, 1)
    md = em.model.model
    assert isinstance(em, expert_model.ExpertModel) and isinstance(md, expert_nn.ScanLSTM)
    F, H = 128, 128
    want = 3 * 4 * F + F * 4 * F + 4 * F + 2 * (F * H + H + H * H + H) + (H * 3 + 3) + (H * 1 + 1)
    assert md.param_count() == want
    flat = torch.from_numpy(synthetic.expert_params_flat(0, md._shapes(), F))
    tree = md.unflatten(flat)
    assert torch.equal(md.flatten(tree), flat)
    hx = torch.randn(4, 3, 3, dtype=torch.float64, generator=torch.Generator().manual_seed(0))
    goal, useq = oexpert.propose(hx, flat.double(), md._shapes(), F, md.head_layers, 5)
    assert goal.shape == (4, 6, 3) and useq.shape == (4, 5, 1) and torch.equal(goal[:, 0], hx[:, -1])
    # the history matters through the LSTM carry only
    hx2 = hx.clone(); hx2[:, 0] += 1.0
    g2, _ = oexpert.propose(hx2, flat.double(), md._shapes(), F, md.head_layers, 5)
    assert not torch.allclose(g2[:, 1:], goal[:, 1:])
    mlp = expert_nn.ScanMLP(3, 16, 3, 1)
    fm = torch.from_numpy(synthetic.expert_params_flat(0, mlp._shapes(), 0)).double()
    g3, _ = oexpert.propose(hx, fm, mlp._shapes(), 0, mlp.head_layers, 5)
    g4, _ = oexpert.propose(hx2, fm, mlp._shapes(), 0, mlp.head_layers, 5)
    assert torch.equal(g3, g4)          # the MLP cell has no carry besides x


def test_bench_work_model_matches_the_survey_table():
    """bench.py's algorithmic FLOPs and HBM bytes per planned state are SURVEY.md 8d's figures (true dims, 1 MAC =
    2 FLOP, forward + input-adjoint backward): the numerator of every roofline fraction in the bench line."""
    import importlib.util
    from tests.conftest import ROOT
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    from gan_mpc_b200 import synthetic
    want_mflop = {"C2": 232.5, "C4": 12046.7, "C5": 572.8}
    for name, mflop in want_mflop.items():
        got = bench.flops_per_state(synthetic.CONFIGS[name]) / 1e6
        assert abs(got - mflop) / mflop < 5e-4, (name, got)
    assert bench.bytes_per_state(synthetic.CONFIGS["C2"]) == 6100     # 3 080 in + 3 020 out
    assert bench.bytes_per_state(synthetic.CONFIGS["C4"]) > 22000
