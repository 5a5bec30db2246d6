"""Multi-agent (MAPPO-Lag) networks on the device -- forward only (first slice).

``MultiAgentNets`` holds the weights of one agent's actor, reward critic and cost critic under the reference's own
``state_dict`` names (safepo/common/model.py:172-363: ``base.feature_norm``, ``base.mlp.fc1``, ``base.mlp.fc2.{i}``,
``act.action_out`` / ``v_out``) and evaluates ``MAPPO_L_Policy.get_actions`` (safepo/multi_agent/mappolag.py:69-82) with
libspo kernels: one ``spo_ma_mlp_layer`` launch per hidden layer and one ``spo_ma_head`` launch per net.  The update of these
nets (mappolag.py:135-199) is not built; there is no CPU path."""
from __future__ import annotations

import ctypes as C

import torch

from safepo import _lib as L


class _Net:
    def __init__(self, state, device, layer_N):
        self.p = {k: v.detach().to(device=device, dtype=torch.float32).contiguous() for k, v in state.items()}
        self.layer_N = layer_N
        self.H = self.p["base.mlp.fc1.0.weight"].shape[0]
        self.D = self.p["base.mlp.fc1.0.weight"].shape[1]

    def features(self, x, work):
        """MLPBase.forward: feature_norm folded into the first layer's launch."""
        p, lib, n = self.p, L.lib(), x.shape[0]
        a, b = work
        L.check(lib.spo_ma_mlp_layer(L.ptr(x), n, self.D, L.ptr(p["base.mlp.fc1.0.weight"]), L.ptr(p["base.mlp.fc1.0.bias"]),
                                     L.ptr(p["base.mlp.fc1.2.weight"]), L.ptr(p["base.mlp.fc1.2.bias"]), self.H,
                                     L.ptr(p["base.feature_norm.weight"]), L.ptr(p["base.feature_norm.bias"]), L.ptr(a), L.stream()),
                "spo_ma_mlp_layer")
        for i in range(self.layer_N):
            L.check(lib.spo_ma_mlp_layer(L.ptr(a), n, self.H, L.ptr(p[f"base.mlp.fc2.{i}.0.weight"]), L.ptr(p[f"base.mlp.fc2.{i}.0.bias"]),
                                         L.ptr(p[f"base.mlp.fc2.{i}.2.weight"]), L.ptr(p[f"base.mlp.fc2.{i}.2.bias"]), self.H,
                                         None, None, L.ptr(b), L.stream()), "spo_ma_mlp_layer")
            a, b = b, a
        return a


class MultiAgentNets:
    def __init__(self, actor_state, critic_state, cost_critic_state, device, layer_N=2, std_x_coef=1.0, std_y_coef=0.5):
        device = torch.device(device)
        if device.type != "cuda":
            raise L.SpoError("MultiAgentNets runs on a CUDA device only (no CPU fallback)")
        self.device = device
        self.actor = _Net(actor_state, device, layer_N)
        self.critic = _Net(critic_state, device, layer_N)
        self.cost_critic = _Net(cost_critic_state, device, layer_N)
        self.act_dim = self.actor.p["act.action_out.fc_mean.weight"].shape[0]
        self.std_x_coef, self.std_y_coef = float(std_x_coef), float(std_y_coef)
        self._work = {}

    def _buffers(self, n, H):
        key = (n, H)
        if key not in self._work:
            self._work[key] = (torch.empty(n, H, dtype=torch.float32, device=self.device), torch.empty(n, H, dtype=torch.float32, device=self.device))
        return self._work[key]

    def _value(self, net, cent_obs):
        n = cent_obs.shape[0]
        feat = net.features(cent_obs, self._buffers(n, net.H))
        out = torch.empty(n, 1, dtype=torch.float32, device=self.device)
        L.check(L.lib().spo_ma_head(L.ptr(feat), n, net.H, L.ptr(net.p["v_out.weight"]), L.ptr(net.p["v_out.bias"]), 1, None, 1.0, 1.0, None,
                                    L.ptr(out), None, L.stream()), "spo_ma_head")
        return out

    def get_actions(self, cent_obs, obs, eps=None, deterministic=False):
        """(values [N,1], actions [N,A], action_log_probs [N,A], cost_preds [N,1]) like MAPPO_L_Policy.get_actions.
        ``eps`` [N,A]: the standard-normal draws to use (torch.randn on the device when omitted and not deterministic)."""
        for t in (cent_obs, obs):
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
                raise L.SpoError("MultiAgentNets needs contiguous fp32 CUDA tensors")
        n, A, net = obs.shape[0], self.act_dim, self.actor
        feat = net.features(obs, self._buffers(n, net.H))
        if deterministic:
            eps = None
        elif eps is None:
            eps = torch.randn(n, A, dtype=torch.float32, device=self.device)
        actions = torch.empty(n, A, dtype=torch.float32, device=self.device)
        logp = torch.empty(n, A, dtype=torch.float32, device=self.device)
        L.check(L.lib().spo_ma_head(L.ptr(feat), n, net.H, L.ptr(net.p["act.action_out.fc_mean.weight"]), L.ptr(net.p["act.action_out.fc_mean.bias"]),
                                    A, L.ptr(net.p["act.action_out.log_std"]), self.std_x_coef, self.std_y_coef, L.ptr(eps), L.ptr(actions),
                                    L.ptr(logp), L.stream()), "spo_ma_head")
        return self._value(self.critic, cent_obs), actions, logp, self._value(self.cost_critic, cent_obs)
