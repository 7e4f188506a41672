"""The stable-baselines-shaped front end on a real GPU: learn / predict / save / load / parameters,
driven the way sb_helper.py:104-128,175 and utils.py:71 drive stable_baselines.SAC."""
import os

import numpy as np
import pytest

import b200grasp
from b200grasp import sb_io
from b200grasp.callbacks import BaseCallback
from oracle import sac_ref as R
from tests.fake_env import FakeGraspEnv
from tests.util import GOLD, load_case

pytestmark = pytest.mark.gpu


class Counter(BaseCallback):
    def __init__(self):
        super().__init__()
        self.steps, self.rollouts = 0, 0

    def _on_step(self):
        self.steps += 1
        return True

    def _on_rollout_end(self):
        self.rollouts += 1


def test_learn_predict_save_load_roundtrip(tmp_path):
    env = b200grasp.VecNormalize(b200grasp.DummyVecEnv([lambda: FakeGraspEnv(1, horizon=15)]), norm_obs=True, norm_reward=True, clip_obs=10.0)
    model = b200grasp.SAC(b200grasp.CnnPolicy, env, policy_kwargs={"layers": [64, 64], "cnn_extractor": None}, verbose=0, gamma=0.99,
                          buffer_size=1000, batch_size=32, learning_rate=3e-4, learning_starts=40, tensorboard_log=None, seed=3)
    cb = Counter()
    model.learn(total_timesteps=120, callback=[cb])
    assert cb.steps == 120 and cb.rollouts == 120
    assert model.n_updates == 120 - 40 + 1 and model.learner.replay_size() == 120
    assert model.get_vec_normalize_env() is env and model.get_env() is env
    obs = env.reset()
    a1, _ = model.predict(obs, deterministic=True)
    assert a1.shape == (1, 5) and np.abs(a1).max() <= 1.0
    params = model.get_parameters()
    assert "model/pi/cnn1/w:0" in params and params["model/pi/cnn1/w:0"].shape == (8, 8, 1, 32)
    path = str(tmp_path / "m" / "sac_model")
    model.save(path)
    env.save(str(tmp_path / "m" / "vecnormalize.pkl"))
    data, zp = sb_io.load_sb_zip(path + ".zip")
    assert data["tau"] == 0.005 and list(zp.keys()) == [k[:-2] for k in params.keys()]
    m2 = b200grasp.SAC.load(path, env)
    a2, _ = m2.predict(obs, deterministic=True)
    assert np.abs(a1 - a2).max() <= 1e-6
    # sb_helper.py:113-115 warm start: load_parameters(get_parameters(), exact_match=False)
    m3 = b200grasp.SAC(b200grasp.CnnPolicy, env, policy_kwargs={"layers": [64, 64], "cnn_extractor": None}, buffer_size=100, batch_size=8)
    m3.load_parameters({k: v for k, v in params.items() if "pi/" in k}, exact_match=False)
    a3, _ = m3.predict(obs, deterministic=True)
    assert np.abs(a1 - a3).max() <= 1e-6
    with pytest.raises(ValueError):
        m3.load_parameters({"model/pi/cnn1/w:0": np.zeros((3, 3))})


def test_load_reference_trained_zip_and_predict(tmp_path):
    """A zip with the reference's exact layout (parameters of trained_models/SAC_depth_1mbuffer/best_model)
    loads, and predict() on the real frame from its vecnormalize.pkl matches the oracle policy."""
    cfg, params, vn = load_case("sac_depth")
    zpath = str(tmp_path / "best_model.zip")
    sb_io.save_sb_zip(zpath, {"gamma": 0.99, "tau": 0.005, "batch_size": 64, "buffer_size": 1000000, "learning_starts": 100,
                              "train_freq": 1, "ent_coef": "auto"}, params)
    model = b200grasp.SAC.load(zpath)
    model.learner.set_norm_stats(vn["obs_mean"], vn["obs_var"], float(vn["ret_var"]), 10.0, 10.0, 1e-8)
    raw = vn["old_obs"].astype(np.float32)
    got = model.learner.act(raw, deterministic=True)
    ref = R.policy_act(params, R.normalize_obs(raw, vn["obs_mean"], vn["obs_var"]), cfg, deterministic=True)
    assert np.abs(got - ref).max() <= 1e-5


def test_train_and_run_cli_on_a_reference_layout_model_dir(tmp_path):
    """f2: the `train` / `run` sub-commands (train_stable_baselines.py:26-109) against the synthetic env: model_dir layout,
    checkpoints, best_model + vecnormalize.pkl, then `run` loads config.yaml / vecnormalize.pkl / the zip and rolls out."""
    import yaml
    from b200grasp import train_cli
    cfg = {"normalize": True, "discount_factor": 0.99, "simplified": False, "reward": {"shaped": False}, "robot": {"discrete": False},
           "simulation": {"real_time": False, "visualize": False},
           "SAC": {"layers": [64, 64], "buffer_size": 2000, "batch_size": 32, "step_size": 3e-4, "total_timesteps": 150, "tensorboard_logs": None}}
    cpath = tmp_path / "cfg.yaml"
    yaml.safe_dump(cfg, open(cpath, "w"))
    mdir = str(tmp_path / "run1")
    model = train_cli.main(["train", "--config", str(cpath), "--algo", "SAC", "--model_dir", mdir, "--env", "tests.fake_env:make_env",
                            "--eval_freq", "60", "--checkpoint_freq", "50", "--n_envs", "2"])
    assert model.n_updates > 0
    for f in ("config.yaml", "best_model/config.yaml", "final_model.zip", "vecnormalize.pkl", "best_model/best_model.zip",
              "best_model/vecnormalize.pkl"):
        assert os.path.exists(os.path.join(mdir, f)), f
    assert any(f.startswith("rl_model_") and f.endswith(".zip") for f in os.listdir(os.path.join(mdir, "logs")))
    assert yaml.safe_load(open(os.path.join(mdir, "config.yaml")))["algorithm"] == "sac"
    model.close()
    out = train_cli.main(["run", "--model", os.path.join(mdir, "best_model", "best_model.zip"), "--env", "tests.fake_env:make_env", "--episodes", "3", "-t"])
    assert out["episodes"] == 3 and np.isfinite(out["mean_reward"]) and out["mean_steps"] == 20
    with pytest.raises(FileExistsError):
        train_cli.main(["train", "--config", str(cpath), "--algo", "SAC", "--model_dir", mdir, "--env", "tests.fake_env:make_env"])
