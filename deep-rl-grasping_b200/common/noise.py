"""``stable_baselines.common.noise`` (sb_helper.py:21): plain host-side noise processes.  ``SAC(action_noise=...)`` itself
is rejected (the shipped configs pass none); the classes exist so that the module imports and DDPG-style code can be read."""
import numpy as np


class AdaptiveParamNoiseSpec:
    def __init__(self, initial_stddev=0.1, desired_action_stddev=0.1, adoption_coefficient=1.01):
        self.initial_stddev, self.desired_action_stddev, self.adoption_coefficient = initial_stddev, desired_action_stddev, adoption_coefficient
        self.current_stddev = initial_stddev

    def adapt(self, distance):
        if distance > self.desired_action_stddev:
            self.current_stddev /= self.adoption_coefficient
        else:
            self.current_stddev *= self.adoption_coefficient


class NormalActionNoise:
    def __init__(self, mean, sigma):
        self._mu, self._sigma = mean, sigma

    def __call__(self):
        return np.random.normal(self._mu, self._sigma)

    def reset(self):
        pass


class OrnsteinUhlenbeckActionNoise:
    def __init__(self, mean, sigma, theta=.15, dt=1e-2, initial_noise=None):
        self._theta, self._mu, self._sigma, self._dt, self.initial_noise = theta, mean, sigma, dt, initial_noise
        self.reset()

    def __call__(self):
        n = self.noise_prev + self._theta * (self._mu - self.noise_prev) * self._dt + \
            self._sigma * np.sqrt(self._dt) * np.random.normal(size=np.shape(self._mu))
        self.noise_prev = n
        return n

    def reset(self):
        self.noise_prev = self.initial_noise if self.initial_noise is not None else np.zeros_like(self._mu)
