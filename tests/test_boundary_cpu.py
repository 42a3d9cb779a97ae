"""Drop-in boundary on CPU: checkpoint resolution and formats (ppo_agent/models.py:54-63, auto_danet.py:161-171),
snapshots (agent.py:245-271), update_model / avg_action / pre_process (agent.py:43-95, 239-243), the ensemble
evaluation loop (eval.py:45-63) and the host-side guards. No CUDA compute: the flat parameter buffers live on the CPU
here, the kernels that consume them are covered by the -m gpu tests."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from cadre_b200 import models as M
from cadre_b200 import ppo_params as P
from cadre_b200._lib import CadreError
from oracle import ref_shim
from oracle import restate as R

needs_ref = pytest.mark.skipif(not ref_shim.available(), reason="/root/reference only exists in the build container")


# ------------------------------------------------------------------------------------------ perception checkpoint
def test_default_checkpoint_path_is_the_references(monkeypatch, tmp_path):
    monkeypatch.setenv("CHALLENGE_DIR", str(tmp_path))
    want = os.path.join(str(tmp_path), "carla_perception/Experiments34", "danet912_nocrash_IL_n10_k1234_r40",
                        "net_epoch90")
    assert M.default_pretrained_path() == want
    cfg = {"vae_params": "CoPM", "measurement_dim": 18}             # the reference's model_cfg has no path key
    obs_dim, vae_params = M.get_vae_output(cfg)
    assert obs_dim == 530 and vae_params.networks["autoencoder"]["pretrained_path"] == want
    assert vae_params.networks["autoencoder"]["z_dims"] == 256
    assert M.get_vae_output({"vae_params": "other", "measurement_dim": 18})[0] == 274   # models.py:41
    monkeypatch.delenv("CHALLENGE_DIR")
    with pytest.raises(KeyError):                                   # os.environ['CHALLENGE_DIR'], auto_danet.py:168
        M.default_pretrained_path()


@needs_ref
def test_default_checkpoint_path_matches_live_reference(tmp_path):
    ref_shim.install(str(tmp_path))
    from carla_perception.Config.auto_danet import danet_config
    assert danet_config().networks["autoencoder"]["pretrained_path"] == M.default_pretrained_path()


def test_checkpoint_file_with_off_path_keys(tmp_path):
    """{'epoch', 'metric', 'autoencoder': state_dict} (experiments_builder.py:442-...): off-path decoder tensors are
    ignored, the folded device weights equal those of the in-memory state, a missing on-path tensor is an error."""
    from cadre_b200.encoder import prepare_weights
    sd = R.danet_fixture_state(0)
    full = dict(sd)
    full["visual_branch.convs.0.weight"] = torch.randn(8, 8, 3, 3)     # decoder heads (danet.py:96-109): off the path
    full["in_bc_speed_fc.0.weight"] = torch.randn(4, 1)
    path = str(tmp_path / "net_epoch90")
    torch.save({"epoch": 90, "metric": {"loss": 0.1}, "autoencoder": full}, path)
    got = M.load_danet_checkpoint(path)
    assert set(got) == set(sd)
    a, b = prepare_weights(got, "cpu"), prepare_weights(sd, "cpu")
    assert all(torch.equal(a[k], b[k]) if torch.is_tensor(a[k]) else a[k] == b[k] for k in b)
    bad = dict(full)
    del bad["backbone.layer3.0.downsample.1.running_var"]
    torch.save({"autoencoder": bad}, path)
    with pytest.raises(CadreError, match="lacks 1 tensor"):
        prepare_weights(M.load_danet_checkpoint(path), "cpu")
    torch.save({"net": full}, path)
    with pytest.raises(CadreError, match="no 'autoencoder' entry"):
        M.load_danet_checkpoint(path)


@needs_ref
def test_checkpoint_written_by_the_reference_network(tmp_path):
    """A full `DANet(cfg).state_dict()` (all decoder heads present) saved the way the reference trainer does."""
    from cadre_b200.encoder import prepare_weights, required_keys
    net, _ = ref_shim.build_reference_danet(R.danet_fixture_state(0))
    state = net.state_dict()
    assert len(state) > len(required_keys())                         # off-path tensors are really there
    path = str(tmp_path / "net_epoch90")
    torch.save({"epoch": 90, "autoencoder": state}, path)
    got = M.load_danet_checkpoint(path)
    ref = prepare_weights(R.danet_fixture_state(0), "cpu")
    mine = prepare_weights(got, "cpu")
    assert all(torch.equal(mine[k], ref[k]) for k in ref if torch.is_tensor(ref[k]))


def test_create_model_refuses_cpu_devices():
    with pytest.raises(CadreError, match="no CPU path"):
        M.create_model({"device_num": -1, "vae_device": 0, "use_lstm": True, "vae_params": "CoPM",
                        "measurement_dim": 18})


def test_fp16_range_check_fails_loudly():
    from cadre_b200.encoder import prepare_weights
    sd = R.danet_fixture_state(0)
    sd["backbone.layer2.0.bn1.running_var"] = torch.full((128,), 1e-14)   # scale = gamma / sqrt(var + 1e-5) ~ 300:
    sd["backbone.layer2.0.conv1.weight"] = sd["backbone.layer2.0.conv1.weight"] * 1e4   # ... x 1e4 weights overflow
    with pytest.raises(CadreError, match="does not fit fp16"):
        prepare_weights(sd, "cpu")
    sd = R.danet_fixture_state(0)
    sd["da_head.conv51.0.weight"] = sd["da_head.conv51.0.weight"] * 1e-5
    with pytest.raises(CadreError, match="below the smallest normal fp16"):
        prepare_weights(sd, "cpu")
    for make in (lambda: R.danet_fixture_state(0, init="xavier"), lambda: R.danet_wide_fixture_state(0)):
        prepare_weights(make(), "cpu")                                  # both robustness fixtures are in range


# ------------------------------------------------------------------------------------------ snapshots
def _flat(seed):
    return M.ModelDict(M.FlatParams("cpu", R.ppo_fixture_state(seed)))


def test_snapshot_round_trip_all_sixteen_modules(tmp_path):
    a, b = _flat(0), _flat(1)
    path = str(tmp_path / "ppo_model_0.pt")
    M.save_model_dict(a, path)
    snap = torch.load(path, weights_only=False)
    assert sorted(snap) == sorted(P.MODULE_ORDER) and len(snap) == 16        # incl. throttle_lstm_* (agent.py:248-258)
    assert not torch.equal(a.owner.params, b.owner.params)
    M.load_model_dict(b, path, torch.device("cpu"))
    assert torch.equal(a.owner.params, b.owner.params)
    # plain dict-of-state-dicts files are accepted too
    torch.save({n: dict(m.state_dict()) for n, m in a.items()}, path)
    c = _flat(2)
    M.load_model_dict(c, path)
    assert torch.equal(a.owner.params, c.owner.params)
    with pytest.raises(ImportError, match="load snapshot error"):            # agent.py:270-271
        M.load_model_dict(c, str(tmp_path / "missing.pt"))


@needs_ref
def test_snapshot_interchange_with_the_reference(tmp_path):
    """(i) a file written by the reference's save_snapshot (pickled nn.Modules, 12 of the 16 modules) loads here;
    (ii) a file written here loads into the reference agent through its own load_snapshot."""
    agent, _ = ref_shim.build_reference_agent(R.danet_fixture_state(0), R.ppo_fixture_state(3))
    path = str(tmp_path / "ref.pt")
    agent.save_snapshot(path)
    mine = _flat(0)
    before = mine.owner.state()
    M.load_model_dict(mine, path)
    after = mine.owner.state()
    ref_sd = R.ppo_fixture_state(3)
    saved = set(torch.load(path, weights_only=False))
    assert len(saved) == 12 and not any(n.startswith("throttle_lstm") for n in saved)
    for m in P.MODULE_ORDER:
        for n in P.module_param_names(m):
            want = ref_sd[m][n] if m in saved else before[m][n]
            assert torch.equal(after[m][n], want), (m, n)
    path2 = str(tmp_path / "mine.pt")
    M.save_model_dict(_flat(5), path2)
    agent.load_snapshot(path2, torch.device("cpu"))
    want = R.ppo_fixture_state(5)
    for m in P.MODULE_ORDER:
        for n, p in agent.model_dict[m].named_parameters():
            assert torch.equal(p.detach(), want[m][n]), (m, n)


# ------------------------------------------------------------------------------------------ agent methods (host)
def _agent_shell(seed=0):
    from cadre_b200.config import load_config
    cfg = load_config().agent_cfg
    md = _flat(seed)
    return SimpleNamespace(model_dict=md, owner=md.owner, STEER_CONTROL=cfg.STEER_CONTROL,
                           THROTTLE_CONTROL=cfg.THROTTLE_CONTROL)


def test_update_model_pulls_parameters():
    from cadre_b200.agent import CadreAgent
    worker, shared = _agent_shell(0), _flat(4)
    CadreAgent.update_model(worker, shared)                                   # flat-to-flat copy (agent.py:239-243)
    assert torch.equal(worker.owner.params, shared.owner.params)
    plain = {n: SimpleNamespace(state_dict=(lambda sd=sd: sd)) for n, sd in R.ppo_fixture_state(6).items()}
    CadreAgent.update_model(worker, plain)                                    # any name -> .state_dict() mapping
    assert torch.equal(worker.owner.params, P.pack_state(R.ppo_fixture_state(6)))
    # module-like surface used by train.py:102 / chief.py:16: named_parameters with .data / .grad views
    names = [n for n, _ in worker.model_dict["steer_lstm_0"].named_parameters()]
    assert names == ["rnn.weight_ih", "rnn.weight_hh", "rnn.bias_ih", "rnn.bias_hh"]


def test_avg_action_and_convert_action():
    from cadre_b200.agent import CadreAgent
    shell = _agent_shell()
    shell.convert_action = lambda a: CadreAgent.convert_action(shell, a)
    t = lambda s, th: [torch.tensor(s), torch.tensor(th)]                     # noqa: E731
    assert CadreAgent.convert_action(shell, t(0, 2)) == [-0.5, 0.6, 0]
    # one agent: the brake is passed through (agent.py:91-94 only touches it for ensembles)
    assert CadreAgent.avg_action(shell, [t(8, 1)]) == [0.0, 0.0, 1.0]
    # three agents, one braking: mean brake 1/3 < 0.5 -> released
    got = CadreAgent.avg_action(shell, [t(8, 1), t(16, 2), t(0, 0)])
    np.testing.assert_allclose(got, [0.0, 0.2, 0.0], atol=1e-12)
    # two of three braking: mean 2/3 stays
    got = CadreAgent.avg_action(shell, [t(8, 1), t(8, 1), t(31, 2)])
    np.testing.assert_allclose(got, [1.0 / 3, 0.2, 2.0 / 3], atol=1e-12)


@needs_ref
def test_avg_action_matches_live_reference():
    from cadre_b200.agent import CadreAgent
    agent, _ = ref_shim.build_reference_agent(R.danet_fixture_state(0), R.ppo_fixture_state(0))
    shell = _agent_shell()
    shell.convert_action = lambda a: CadreAgent.convert_action(shell, a)
    rs = np.random.RandomState(0)
    for n in (1, 2, 3, 5):
        acts = [[torch.tensor(int(rs.randint(33))), torch.tensor(int(rs.randint(3)))] for _ in range(n)]
        assert CadreAgent.avg_action(shell, acts) == agent.avg_action(acts)


def test_pre_process_matches_oracle_bit_for_bit():
    from cadre_b200.agent import pre_process_tensors
    rs = np.random.RandomState(2000)
    tick = R.synthetic_tick(rs)
    tick["route_fig"][3] = (rs.rand(256, 144) * 200).astype(np.uint8)   # non-binary: uint8 truncation quirk
    tick["route_fig"][5] = 0                                             # max == 0 branch
    keep = tick["route_fig"].copy()
    got = pre_process_tensors(torch.from_numpy(tick["rgb"]), torch.from_numpy(tick["route_fig"]))
    assert np.array_equal(tick["route_fig"], keep)                       # caller's array untouched
    want = R.pre_process(tick["rgb"], tick["route_fig"].copy())
    assert got.shape == (8, 4, 144, 256) and got.dtype == torch.float32
    assert np.array_equal(got.numpy(), want)


# ------------------------------------------------------------------------------------------ learner host logic
def test_ragged_minibatch_is_rejected():
    from cadre_b200.learner import Learner
    from cadre_b200.storage import RolloutStorage
    L = Learner.__new__(Learner)
    L.mini_batch, L._rng_states = 100, []
    ok = [tuple(RolloutStorage(200, 2, 8, 1, 8, True, 0.99, 0.95) for _ in range(2))]
    assert L.sample_epoch_indices(ok).shape == (2, 1, 2, 100)
    ragged = [tuple(RolloutStorage(201, 2, 8, 1, 8, True, 0.99, 0.95) for _ in range(2))]   # chunks 100, 100, 1
    with pytest.raises(CadreError, match="ragged minibatch"):
        L.sample_epoch_indices(ragged)


# ------------------------------------------------------------------------------------------ evaluation ensemble
class _ScriptedAgent:
    """Acts like CadreAgent.act's 5-tuple with a fixed action; avg_action is the product's."""

    def __init__(self, steer, throttle):
        from cadre_b200.agent import CadreAgent
        shell = _agent_shell()
        self.STEER_CONTROL, self.THROTTLE_CONTROL = shell.STEER_CONTROL, shell.THROTTLE_CONTROL
        self._a = [torch.tensor(steer), torch.tensor(throttle)]
        self._cls = CadreAgent
        self.loaded = None

    def act(self, tick):
        assert tick["rgb"].shape == (8, 144, 256, 3)
        return None, self._a, None, None, None

    def convert_action(self, a):
        return self._cls.convert_action(self, a)

    def avg_action(self, lst):
        return self._cls.avg_action(self, lst)

    def load_snapshot(self, path, device):
        self.loaded = path


def test_eval_ensemble_loop(tmp_path):
    from cadre_b200 import eval as E
    from cadre_b200.synthetic_env import SyntheticEnv
    made = []

    def make():
        made.append(_ScriptedAgent((8, 16, 31)[len(made)], 1 if len(made) == 0 else 2))
        return made[-1]
    group = E.load_agent_group(make, str(tmp_path), [100, 200, 300])
    assert [a.loaded for a in group] == [os.path.join(str(tmp_path), "models", f"ppo_model_{e}.pt")
                                         for e in (100, 200, 300)]
    env = SyntheticEnv(dict(seed=3, done_prob=0.2))
    res = E.evaluate(group, env, eval_episode=3)
    assert len(res) == 3 and all(r["steps"] >= 1 for r in res)
    # steer 0, 8/16, 1 -> mean 0.5; throttle (0, 0.6, 0.6)/3 = 0.4; brake (1, 0, 0)/3 < 0.5 -> released
    for r in res:
        for c in r["controls"]:
            np.testing.assert_allclose(c, [0.5, 0.4, 0.0], atol=1e-12)
    with pytest.raises(ValueError):
        E.evaluate([], env, 1)
