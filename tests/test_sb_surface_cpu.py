"""The import surface and host-side helpers a maintainer gets after ``s/stable_baselines/b200grasp/`` in the reference's
training harness (sb_helper.py:6-21, base_callbacks.py:11-14, train_stable_baselines.py:7-18).  CPU only."""
import importlib
import json
import os

import numpy as np
import pytest

import b200grasp as sb
from b200grasp.bench import Monitor
from b200grasp.bench.monitor import load_results
from b200grasp.common.callbacks import CheckpointCallback, EvalCallback, EveryNTimesteps
from b200grasp.common.evaluation import evaluate_policy
from b200grasp.common.vec_env import DummyVecEnv, VecEnv, VecNormalize, sync_envs_normalization
from tests.fake_env import FakeGraspEnv

# (module under stable_baselines, names) exactly as the reference imports them
REFERENCE_IMPORTS = [
    ("common.policies", ["MlpPolicy", "CnnPolicy"]),                                            # sb_helper.py:11,18
    ("deepq.policies", ["MlpPolicy"]),                                                          # sb_helper.py:12
    ("common", ["set_global_seeds"]),                                                           # sb_helper.py:13
    ("sac.policies", ["MlpPolicy", "CnnPolicy", "LnCnnPolicy"]),                                # sb_helper.py:15-17
    ("common.vec_env", ["DummyVecEnv", "SubprocVecEnv", "VecNormalize", "VecFrameStack", "VecEnv", "sync_envs_normalization"]),
    ("common.callbacks", ["BaseCallback", "CheckpointCallback", "EvalCallback", "EventCallback"]),   # sb_helper.py:20, base_callbacks.py:14
    ("common.noise", ["NormalActionNoise", "OrnsteinUhlenbeckActionNoise", "AdaptiveParamNoiseSpec"]),
    ("common.evaluation", ["evaluate_policy"]),                                                 # base_callbacks.py:12
    ("bench", ["Monitor"]),                                                                     # train_stable_baselines.py:18
    ("", ["SAC", "BDQ", "logger"]),                                                             # train_stable_baselines.py:7,12
]


@pytest.mark.parametrize("mod,names", REFERENCE_IMPORTS)
def test_every_name_the_reference_imports_exists(mod, names):
    m = importlib.import_module("b200grasp" + ("." + mod if mod else ""))
    for n in names:
        assert hasattr(m, n), (mod, n)


def test_out_of_scope_algorithms_and_policies_fail_loudly():
    for algo in ("DQN", "TRPO", "PPO2", "DDPG"):
        with pytest.raises(NotImplementedError):
            getattr(sb, algo)
    from b200grasp.sac.policies import LnCnnPolicy
    with pytest.raises(NotImplementedError):
        sb.SAC(LnCnnPolicy, None, _init_setup_model=False)
    from b200grasp.common.vec_env import VecFrameStack
    with pytest.raises(NotImplementedError):
        VecFrameStack([])


class _ConstModel:
    """predict() -> fixed action; save() records the path (what the callbacks need from a model)."""

    def __init__(self, env, n_act):
        self.env, self.n_act, self.saved, self.num_timesteps = env, n_act, [], 0

    def predict(self, obs, state=None, mask=None, deterministic=True):
        obs = np.asarray(obs)
        batched = obs.ndim == len(self.env.observation_space.shape) + 1
        a = np.zeros((obs.shape[0], self.n_act), np.float32) if batched else np.zeros(self.n_act, np.float32)
        return a, None

    def save(self, path):
        self.saved.append(path)

    def get_env(self):
        return self.env


def test_monitor_writes_the_shipped_column_layout(tmp_path):
    env = Monitor(FakeGraspEnv(seed=1, horizon=5, obs_shape=(4, 4, 2)), str(tmp_path / "log_file"))
    for _ in range(3):
        env.reset()
        done = False
        while not done:
            _, _, done, info = env.step(np.zeros(5, np.float32))
        assert "episode" in info
    env.close()
    lines = open(tmp_path / "log_file.monitor.csv").read().splitlines()
    head = json.loads(lines[0][1:])
    assert set(head) == {"t_start", "env_id"}                 # '#{"t_start": ..., "env_id": ...}' as in trained_models/*/log_file.monitor.csv
    assert lines[1] == "r,s,l,c,timesteps,t"
    rows = [l.split(",") for l in lines[2:]]
    assert [int(r[2]) for r in rows] == [5, 5, 5] and [int(r[4]) for r in rows] == [4, 9, 14]      # timesteps = total - 1 (149, 299, ... shipped)
    assert env.get_episode_lengths() == [5, 5, 5] and len(load_results(str(tmp_path))) == 3
    with pytest.raises(RuntimeError):
        env.step(np.zeros(5, np.float32))                     # needs reset


def test_evaluate_policy_and_sync_envs_normalization():
    train = VecNormalize(DummyVecEnv([lambda: FakeGraspEnv(seed=2, horizon=7, obs_shape=(4, 4, 2))]))
    evalv = VecNormalize(DummyVecEnv([lambda: FakeGraspEnv(seed=3, horizon=7, obs_shape=(4, 4, 2))]), training=False)
    assert isinstance(train, VecEnv)
    obs = train.reset()
    for _ in range(20):
        obs, *_ = train.step(np.zeros((1, 5), np.float32))
    assert not np.allclose(train.obs_rms.mean, evalv.obs_rms.mean)
    sync_envs_normalization(train, evalv)
    assert np.array_equal(train.obs_rms.mean, evalv.obs_rms.mean) and train.obs_rms is not evalv.obs_rms
    assert np.array_equal(train.ret_rms.var, evalv.ret_rms.var)
    model = _ConstModel(evalv, 5)
    rews, lens = evaluate_policy(model, evalv, n_eval_episodes=3, return_episode_rewards=True)
    assert lens == [7, 7, 7] and len(rews) == 3
    mean, std = evaluate_policy(model, evalv, n_eval_episodes=2)
    assert np.isfinite(mean) and std >= 0
    with pytest.raises(AssertionError):
        evaluate_policy(model, evalv, n_eval_episodes=1, reward_threshold=1e9)


def test_checkpoint_eval_and_every_n_callbacks(tmp_path):
    train = DummyVecEnv([lambda: FakeGraspEnv(seed=4, horizon=4, obs_shape=(4, 4, 2))])
    model = _ConstModel(train, 5)
    ck = CheckpointCallback(save_freq=3, save_path=str(tmp_path / "ck"), name_prefix="rl_model")
    fired = []

    class Mark(sb.callbacks.BaseCallback):
        def _on_step(self):
            fired.append(self.num_timesteps)
            return True

    ev = EvalCallback(FakeGraspEnv(seed=5, horizon=4, obs_shape=(4, 4, 2)), callback_on_new_best=Mark(), n_eval_episodes=2, eval_freq=5,
                      log_path=str(tmp_path / "ev"), best_model_save_path=str(tmp_path / "best"), verbose=0)
    every = EveryNTimesteps(4, Mark())
    for cb in (ck, ev, every):
        cb.init_callback(model)
        cb.on_training_start({}, {})
    for t in range(1, 11):
        model.num_timesteps = t
        for cb in (ck, ev, every):
            assert cb.on_step()
    assert [os.path.basename(p) for p in model.saved if "rl_model" in p] == ["rl_model_3_steps", "rl_model_6_steps", "rl_model_9_steps"]
    assert any(p.endswith("best_model") for p in model.saved)           # first evaluation is always a new best
    z = np.load(str(tmp_path / "ev" / "evaluations.npz"))
    assert list(z["timesteps"]) == [5, 10] and z["results"].shape == (2, 2)
    assert 5 in fired and 4 in fired and 8 in fired


def test_subproc_vec_env_steps_environments_in_worker_processes():
    """The host actor loop of BASELINE config 5 (sb_helper.py:19 imports SubprocVecEnv): N environments in worker
    processes, auto-reset with the terminal observation kept in info, same arrays as DummyVecEnv on the same seeds."""
    from b200grasp.vec_env import DummyVecEnv, SubprocVecEnv
    from tests.fake_env import FakeGraspEnv
    fns = [(lambda i=i: FakeGraspEnv(seed=i, horizon=3)) for i in range(3)]
    sub, dum = SubprocVecEnv(fns), DummyVecEnv(fns)
    try:
        assert sub.num_envs == 3 and tuple(sub.observation_space.shape) == (64, 64, 2)
        o1, o2 = sub.reset(), dum.reset()
        assert o1.shape == (3, 64, 64, 2) and np.array_equal(o1, o2)
        for t in range(4):
            a = np.full((3, 5), 0.6 if t % 2 else -0.6, np.float32)
            r1, r2 = sub.step(a), dum.step(a)
            assert np.array_equal(r1[0], r2[0]) and np.array_equal(r1[1], r2[1]) and np.array_equal(r1[2], r2[2])
            if t == 2:
                assert r1[2].all() and all("terminal_observation" in i for i in r1[3])
        assert sub.get_attr("horizon") == [3, 3, 3]
    finally:
        sub.close()


def test_cnn_policy_without_extractor_is_refused_and_cli_parses():
    """CnnPolicy with policy_kwargs={} is stable-baselines' plain nature_cnn (sb_helper.py:93-95), not the
    augmented extractor -- refuse instead of silently building a different network.  Also the train / run CLI surface."""
    from b200grasp.vec_env import DummyVecEnv
    from b200grasp.train_cli import build_parser
    from tests.fake_env import FakeGraspEnv
    env = DummyVecEnv([lambda: FakeGraspEnv()])
    with pytest.raises(NotImplementedError, match="nature_cnn"):
        sb.SAC(sb.CnnPolicy, env, policy_kwargs={})
    p = build_parser()
    a = p.parse_args(["train", "--config", "c.yaml", "--algo", "SAC", "--model_dir", "out", "-s", "-sh", "--timestep", "1000"])
    assert a.func.__name__ == "train" and a.simple and a.shaped and a.timestep == "1000" and a.load_dir is None
    r = p.parse_args(["run", "--model", "m/best_model.zip", "-t", "-s"])
    assert r.func.__name__ == "run" and r.test and r.stochastic
