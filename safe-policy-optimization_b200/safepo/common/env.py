"""Environment factory of the single-agent trainers (reference safepo/common/env.py:35-80).

``make_sa_mujoco_env(num_envs, env_id, seed)`` returns ``(env, obs_space, act_space)`` like the
reference's factory.  The MuJoCo simulation itself stays on host cores (north star: "env.step stays
on host"); what moves is the arithmetic the reference wraps around it:

* ``SafeRescaleAction(env, -1, 1)`` (env.py:62,76)   -> ``spo_action_rescale`` on the device,
* ``SafeNormalizeObservation(env)`` (env.py:66,77)  -> ``spo_obs_normalize`` on the device,

both applied by ``_engine.Rollout`` right before / after the PCIe copies of a step
(``--normalize-obs``, ``args.rescale_action``), so the host env handed back here is the *bare*
vector env.  ``safety_gymnasium`` is a third-party package that is not part of this repository
(and not installable offline): when it is missing the factory raises :class:`SpoError` with that
message instead of an ImportError from deep inside a trainer -- there is no fallback env; the
synthetic stream is selected explicitly with ``--env synthetic``.
"""
from __future__ import annotations

from safepo import _lib as L


def _require_safety_gymnasium():
    try:
        import safety_gymnasium  # noqa: F401
        return safety_gymnasium
    except ImportError as e:  # pragma: no cover - depends on the host
        raise L.SpoError("--env mujoco needs the safety_gymnasium package on the host "
                         "(pip install safety-gymnasium); it is not bundled with libspo.  "
                         "Use --env synthetic for the observation-shaped synthetic stream.") from e


class HostVectorEnv:
    """The bare host vector env plus the two attributes the trainers read off it.

    ``device_wrappers`` tells ``Rollout`` which of the reference's wrappers it has to apply on the
    device for this env (the reference applies both, env.py:62-66)."""

    device_wrappers = ("rescale_action", "normalize_obs")

    def __init__(self, venv, obs_space, act_space):
        self.venv = venv
        self.observation_space = obs_space
        self.action_space = act_space
        self.single_observation_space = obs_space
        self.single_action_space = act_space
        self.obs_rms = None          # filled by Rollout with the device-side RunningMeanStd (checkpoint "Normalizer")

    def reset(self, seed=None):
        return self.venv.reset(seed=seed)

    def step(self, action):
        return self.venv.step(action)

    def close(self):
        return self.venv.close()


def make_sa_mujoco_env(num_envs: int, env_id: str, seed: int | None = None):
    """Same signature and return triple as reference env.py:35.  num_envs == 1 is vectorised too
    (the reference's SafeUnsqueeze gives the same [1, D] shapes)."""
    sg = _require_safety_gymnasium()
    from safety_gymnasium.vector.async_vector_env import SafetyAsyncVectorEnv
    from safety_gymnasium.vector.sync_vector_env import SafetySyncVectorEnv

    def create_env():
        return sg.make(env_id)

    fns = [create_env for _ in range(num_envs)]
    venv = SafetyAsyncVectorEnv(fns) if num_envs > 1 else SafetySyncVectorEnv(fns)
    venv.reset(seed=seed)
    env = HostVectorEnv(venv, venv.single_observation_space, venv.single_action_space)
    return env, env.observation_space, env.action_space
