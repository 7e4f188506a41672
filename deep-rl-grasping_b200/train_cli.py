"""``train`` / ``run`` command line of the reference, re-hosted on the B200 learner.

Mirrors /root/reference/manipulation_main/training/train_stable_baselines.py:26-148 (same sub-commands, flags,
``model_dir`` layout: ``config.yaml``, ``best_model/``, ``logs/rl_model_*`` checkpoints, ``vecnormalize.pkl``,
``log_file.monitor.csv``) and the SAC / BDQ branches of ``SBPolicy.learn`` (sb_helper.py:69-128,175-247).  The
environment itself stays the reference's (PyBullet on host cores): ``--env module:callable`` names a factory
``f(config, evaluate=False, validate=False, test=False) -> gym.Env``; the default imports the reference package and
calls ``gym.make('gripper-env-v0', ...)`` exactly like the original script.

  python -m b200grasp.train_cli train --config config/gripper_grasp.yaml --algo SAC --model_dir out/sac_depth
  python -m b200grasp.train_cli run --model trained_models/SAC_depth_1mbuffer/best_model/best_model.zip -t
"""
from __future__ import annotations

import argparse
import importlib
import logging
import os

import numpy as np
import yaml

from . import BDQ, SAC
from .bench import Monitor
from .callbacks import BaseCallback, CheckpointCallback, EvalCallback
from .sac_model import CnnPolicy, MlpPolicy
from .vec_env import DummyVecEnv, SubprocVecEnv, VecNormalize


def _default_factory(config, evaluate=False, validate=False, test=False):
    import gym                      # the reference's own dependency chain (gym + pybullet + manipulation_main)
    import manipulation_main        # noqa: F401  (registers gripper-env-v0)
    return gym.make("gripper-env-v0", config=config, evaluate=evaluate, validate=validate, test=test)


def _factory(spec):
    if not spec:
        return _default_factory
    mod, _, fn = spec.partition(":")
    return getattr(importlib.import_module(mod), fn)


class SaveVecNormalizeCallback(BaseCallback):
    """sb_helper.py:24-56: writes ``vecnormalize.pkl`` next to the (best) model."""

    def __init__(self, save_freq, save_path, name_prefix=None, verbose=0):
        super().__init__(verbose)
        self.save_freq, self.save_path, self.name_prefix = save_freq, save_path, name_prefix

    def _init_callback(self):
        os.makedirs(self.save_path, exist_ok=True)

    def _on_step(self):
        if self.n_calls % self.save_freq == 0:
            name = "vecnormalize.pkl" if self.name_prefix is None else f"{self.name_prefix}_{self.num_timesteps}_steps.pkl"
            vn = self.model.get_vec_normalize_env()
            if vn is not None:
                vn.save(os.path.join(self.save_path, name))
        return True


def _is_image_obs(env):
    return len(env.observation_space.shape) == 3


def train(args):
    config = yaml.safe_load(open(args.config))
    os.mkdir(args.model_dir)                                   # like the reference: refuses to overwrite a run
    os.mkdir(os.path.join(args.model_dir, "best_model"))
    algo = args.algo
    if args.simple:
        config["simplified"] = True
    if args.shaped:
        config["reward"]["shaped"] = True
    if args.timestep:
        config[algo]["total_timesteps"] = int(args.timestep)
    config["robot"]["discrete"] = algo == "DQN"
    config[algo]["save_dir"] = args.model_dir
    config["algorithm"] = algo.lower()
    make = _factory(args.env)
    n_envs = max(1, int(args.n_envs))
    if n_envs == 1:
        env = DummyVecEnv([lambda: Monitor(make(config), os.path.join(args.model_dir, "log_file"))])
    else:                                                      # BASELINE config 5: vectorised host actor loop
        env = SubprocVecEnv([(lambda i=i: Monitor(make(config), os.path.join(args.model_dir, f"log_file_{i}"))) for i in range(n_envs)])
    for sub in ("", "best_model"):
        yaml.safe_dump(config, open(os.path.join(args.model_dir, sub, "config.yaml"), "w"))
    test_env = DummyVecEnv([lambda: make(config, evaluate=True, validate=True)])
    norm = bool(config.get("normalize", False))
    eval_path = os.path.join(args.model_dir, "best_model")
    if norm:
        test_env = VecNormalize(test_env, norm_obs=True, norm_reward=False, clip_obs=10.0)
    callbacks = [
        EvalCallback(test_env, best_model_save_path=eval_path, log_path=os.path.join(eval_path, "logs"), eval_freq=args.eval_freq,
                     n_eval_episodes=10, callback_on_new_best=SaveVecNormalizeCallback(1, eval_path), deterministic=True),
        CheckpointCallback(save_freq=args.checkpoint_freq, save_path=os.path.join(args.model_dir, "logs"), name_prefix="rl_model"),
    ]
    top = os.path.dirname(args.load_dir) if args.load_dir else None
    if norm:
        if args.load_dir:
            env = VecNormalize.load(os.path.join(top, "vecnormalize.pkl"), VecNormalize(env, training=True, norm_obs=False, norm_reward=False, clip_obs=10.0))
        else:
            env = VecNormalize(env, norm_obs=True, norm_reward=True, clip_obs=10.0)
    c = config[algo]
    if algo == "SAC":
        if _is_image_obs(env):
            policy, kw = CnnPolicy, {"layers": c["layers"], "cnn_extractor": "augmented_nature_cnn"}
        else:
            policy, kw = MlpPolicy, {"layers": c["layers"], "layer_norm": False}
        model = SAC(policy, env, policy_kwargs=kw, verbose=1, gamma=config["discount_factor"], buffer_size=c["buffer_size"],
                    batch_size=c["batch_size"], learning_rate=c["step_size"], precision=args.precision)
        if args.load_dir:
            old = SAC.load(args.load_dir, env, buffer_size=1)
            model.load_parameters(old.get_parameters(), exact_match=False)
            old.close()
    elif algo == "BDQ":
        model = BDQ("MlpActPolicy", env, policy_kwargs={"layers": c["layers"]}, verbose=1, gamma=config["discount_factor"],
                    batch_size=c["batch_size"], buffer_size=c["buffer_size"], learning_rate=c["step_size"],
                    exploration_fraction=c.get("exploration_fraction", 0.1), exploration_final_eps=c.get("exploration_final_eps", 0.02),
                    num_actions_pad=c.get("num_actions_pad", 33), learning_starts=c.get("learning_starts", 1000),
                    target_network_update_freq=c.get("target_network_update_freq", 1000),
                    prioritized_replay=c.get("prioritized_replay", False))
        if args.load_dir:
            model.load_parameters(BDQ.load(args.load_dir, env).get_parameters())
    else:
        raise NotImplementedError(f"--algo {algo}: the B200 learner builds the SAC and BDQ branches of SBPolicy.learn (sb_helper.py:85-226)")
    model.learn(total_timesteps=int(c["total_timesteps"]), callback=callbacks)
    model.save(os.path.join(args.model_dir, "final_model" if algo != "BDQ" else "bdq_model"))   # sb_helper.py:228-247
    vn = model.get_vec_normalize_env()
    if vn is not None:
        vn.save(os.path.join(args.model_dir, "vecnormalize.pkl"))
    env.close()
    test_env.close()
    return model


def run_agent(task, agent, stochastic=False, n_episodes=100):
    """manipulation_main/utils.py:14-76: roll out `n_episodes` and report success rate / reward / length."""
    rewards, steps, successes = [], [], []
    for _ in range(n_episodes):
        obs, done = task.reset(), np.array([False])
        ep_r, ep_n, info = 0.0, 0, {}
        while not done[0]:
            action = agent.predict(obs, deterministic=not stochastic)[0]
            obs, r, done, infos = task.step(action)
            rew = task.get_original_reward() if hasattr(task, "get_original_reward") else r
            ep_r += float(np.ravel(rew)[0]); ep_n += 1
            info = infos[0] if infos else {}
        rewards.append(ep_r); steps.append(ep_n)
        successes.append(bool(info.get("is_success", info.get("status", None) in ("SUCCESS", 1))))
    out = {"success_rate": float(np.mean(successes)), "mean_reward": float(np.mean(rewards)), "mean_steps": float(np.mean(steps)), "episodes": n_episodes}
    print(f"Finished {n_episodes} episodes: success rate {out['success_rate']:.3f}, mean reward {out['mean_reward']:.1f}, mean steps {out['mean_steps']:.1f}")
    return out


def run(args):
    top = os.path.dirname(args.model)
    config = yaml.safe_load(open(os.path.join(top, "config.yaml")))
    make = _factory(args.env)
    task = DummyVecEnv([lambda: make(config, evaluate=True, test=args.test)])
    if config.get("normalize", False):
        task = VecNormalize.load(os.path.join(top, "vecnormalize.pkl"), VecNormalize(task, training=False, norm_obs=True, norm_reward=True, clip_obs=10.0))
        task.training = False
    algo = config["algorithm"]
    if algo == "sac":
        agent = SAC.load(args.model, precision=args.precision)
        agent._vec_normalize_env = task if isinstance(task, VecNormalize) else None     # predict() receives normalised observations
        if agent._vec_normalize_env is not None:
            agent._sync_norm_stats()
    elif algo == "bdq":
        agent = BDQ.load(args.model)
    else:
        raise NotImplementedError(f"algorithm '{algo}': only sac / bdq zips run on the B200 learner")
    print("Run the agent")
    out = run_agent(task, agent, args.stochastic, n_episodes=args.episodes)
    task.close()
    return out


def build_parser():
    p = argparse.ArgumentParser(prog="b200grasp.train_cli")
    sub = p.add_subparsers()
    t = sub.add_parser("train")
    t.add_argument("--config", type=str, required=True)
    t.add_argument("--algo", type=str, required=True)
    t.add_argument("--model_dir", type=str, required=True)
    t.add_argument("--load_dir", type=str)
    t.add_argument("--timestep", type=str)
    t.add_argument("-s", "--simple", action="store_true")
    t.add_argument("-sh", "--shaped", action="store_true")
    t.add_argument("-v", "--visualize", action="store_true")
    t.add_argument("-tf", "--timefeature", action="store_true")
    t.add_argument("--env", type=str, default=None, help="module:callable environment factory (default: the reference's gripper-env-v0)")
    t.add_argument("--n_envs", type=int, default=1, help=">1: SubprocVecEnv actor loop on host cores feeding the device replay")
    t.add_argument("--precision", default="bf16x3", choices=["fp32", "bf16x3", "bf16"])
    t.add_argument("--eval_freq", type=int, default=50000)
    t.add_argument("--checkpoint_freq", type=int, default=25000)
    t.set_defaults(func=train)
    r = sub.add_parser("run")
    r.add_argument("--model", type=str, required=True)
    r.add_argument("-v", "--visualize", action="store_true")
    r.add_argument("-t", "--test", action="store_true")
    r.add_argument("-s", "--stochastic", action="store_true")
    r.add_argument("--env", type=str, default=None)
    r.add_argument("--episodes", type=int, default=100)
    r.add_argument("--precision", default="bf16x3", choices=["fp32", "bf16x3", "bf16"])
    r.set_defaults(func=run)
    return p


def main(argv=None):
    logging.getLogger().setLevel(logging.INFO)
    args = build_parser().parse_args(argv)
    if not hasattr(args, "func"):
        build_parser().print_help()
        return None
    return args.func(args)


if __name__ == "__main__":
    main()
