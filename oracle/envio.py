"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): CPU restatement of the two env wrappers
between ``env.step`` and ``policy.step``.

PARITY UNPINNED: the arithmetic lives in third-party packages that are absent from
/root/reference and from this image (gymnasium -- required through safety-gymnasium,
un-pinned in the reference's setup.py:32).  What follows restates the published algorithm of
gymnasium 0.28/0.29 ``wrappers/normalize.py`` (RunningMeanStd, update_mean_var_count_from_moments,
NormalizeObservation.normalize) and ``wrappers/rescale_action.py``; the anchors on the reference
side are its call sites: safepo/common/wrappers.py:42-49 (SafeNormalizeObservation.step ->
self.normalize(obs)), safepo/common/env.py:62,66,76-77 (SafeRescaleAction(env, -1.0, 1.0),
SafeNormalizeObservation(env)) and ppo_lag.py:381-386 (env.obs_rms checkpointed as "Normalizer")."""
import numpy as np


class RunningMeanStd:
    def __init__(self, epsilon=1e-4, shape=()):
        self.mean = np.zeros(shape, "float64")
        self.var = np.ones(shape, "float64")
        self.count = epsilon

    def update(self, x):
        batch_mean = np.mean(x, axis=0)
        batch_var = np.var(x, axis=0)
        batch_count = x.shape[0]
        delta = batch_mean - self.mean
        tot_count = self.count + batch_count
        new_mean = self.mean + delta * batch_count / tot_count
        m_a = self.var * self.count
        m_b = batch_var * batch_count
        M2 = m_a + m_b + np.square(delta) * self.count * batch_count / tot_count
        self.mean, self.var, self.count = new_mean, M2 / tot_count, tot_count


class NormalizeObservation:
    def __init__(self, obs_dim, epsilon=1e-8):
        self.obs_rms = RunningMeanStd(shape=(obs_dim,))
        self.epsilon = epsilon

    def normalize(self, obs, update=True):
        if update:
            self.obs_rms.update(obs)
        return (obs - self.obs_rms.mean) / np.sqrt(self.obs_rms.var + self.epsilon)


def rescale_action(action, low, high, min_action=-1.0, max_action=1.0):
    action = np.clip(action, min_action, max_action)
    action = low + (high - low) * ((action - min_action) / (max_action - min_action))
    return np.clip(action, low, high)
