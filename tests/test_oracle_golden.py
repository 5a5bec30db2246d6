"""Pins oracle/ against fixtures produced by the real reference (tests/golden/make_golden.py).
Bit-exact comparisons: the oracle runs the same torch-CPU ops in the same order."""
import numpy as np
import pytest
import torch

from oracle import spo_oracle as O
from oracle import trainers as TR


def _policy_from(state, D, A):
    pol = O.OraclePolicy(D, A, [64, 64])
    pol.load(state)
    return pol


def test_forward_bit_exact(golden):
    for c in golden("forward")["forward"]:
        pol = _policy_from(c["state"], c["D"], c["A"])
        with torch.no_grad():
            act, logp, vr, vc = O.policy_step(pol, c["obs"], eps=c["eps"])
            dact, dlogp, _, _ = O.policy_step(pol, c["obs"], deterministic=True)
            a1, l1, r1, _ = O.policy_step(pol, c["obs"][0], eps=c["eps"][0])
        for got, want in ((act, c["act"]), (logp, c["logp"]), (vr, c["v_r"]), (vc, c["v_c"]), (dact, c["det_act"]),
                          (dlogp, c["det_logp"])):
            assert torch.equal(got, want)
        assert [tuple(a1.shape), tuple(l1.shape), tuple(r1.shape)] == [tuple(s) for s in c["row_shapes"]]


def test_init_rng_order_matches_reference(golden):
    c = golden("forward")["forward"][0]
    torch.manual_seed(0)
    pol = O.OraclePolicy(c["D"], c["A"], [64, 64])
    for net in O.NET_ORDER:
        for k, v in c["state"][net].items():
            if k == "log_std":
                continue
            assert torch.equal(pol.nets[net][k].detach(), v), (net, k)


def test_gae_known_answers(golden):
    g = golden("gae")
    adv, tgt = O.gae_path(torch.tensor([0.5, 0.4, 0.3, 0.2]), torch.tensor([1.0, 0.0, 2.0, 0.2]), 0.95, 0.99)
    assert torch.equal(adv, g["gae_kat"]["adv"]) and torch.equal(tgt, g["gae_kat"]["tgt"])
    # SURVEY Appendix B.1 literal values
    assert adv.tolist() == [2.4779859377461078, 1.6820690165758132, 1.8980000019073486]


def test_gae_dual_and_finalize_vs_reference_buffer(golden):
    for c in golden("gae")["gae_cases"]:
        outs = O.gae_dual(c["rew"], c["cost"], c["v_r"], c["v_c"], c["seg_end"], c["boot_r"], c["boot_c"],
                          c["gamma"], c["lam"], c["lam_c"])
        for got, key in zip(outs, ("adv_r", "adv_c", "target_value_r", "target_value_c")):
            assert torch.equal(got, c["raw"][key]), key
        a, cc, _ = O.adv_finalize(outs[0].reshape(-1), outs[1].reshape(-1), 0.0)
        assert torch.equal(a, c["get"]["adv_r"]) and torch.equal(cc, c["get"]["adv_c"])


def test_buffer_layout_known_answer(golden):
    kat = golden("gae")["buffer_kat"]
    assert kat["obs"][:, 0].tolist() == [0, 1, 2, 3, 100, 101, 102, 103]  # env-major, SURVEY fact 4
    buf = TR.PathBuffer(2, 4, 1, 1, 0.99)
    for t in range(4):
        z = torch.tensor([float(t), 10.0 + t])
        buf.store(t, torch.tensor([[t + 0.0], [t + 100.0]]), torch.zeros(2, 1), z, z / 2, z / 10, z / 5, torch.zeros(2))
        if t == 1:
            buf.finish_path(torch.zeros(1), torch.zeros(1), 0)
        if t == 3:
            buf.finish_path(torch.tensor([0.7]), torch.tensor([0.3]), 0)
            buf.finish_path(torch.tensor([1.1]), torch.tensor([0.9]), 1)
    data = buf.get()
    for k in ("obs", "adv_r", "adv_c", "target_value_r", "target_value_c"):
        assert torch.equal(data[k], kat[k]), k


def test_lagrange(golden):
    g = golden("lagrange")["lagrange"]
    L = O.OracleLagrange(25.0, 0.001, 0.035)
    got = []
    for jc in g["jc"]:
        L.update_lagrange_multiplier(jc)
        got.append(L.lagrangian_multiplier)
    assert got == g["lam"]
    assert got[:4] == [0.0, 0.0018421054119244218, 0.015919595956802368, 0.020150989294052124]  # Appendix B.3
    L2 = O.OracleLagrange(25.0, 0.001, 0.035, upper_bound=2.0)
    got2 = []
    for _ in g["capped"]:
        L2.update_lagrange_multiplier(500.0)
        got2.append(L2.lagrangian_multiplier)
    assert got2 == g["capped"] and max(got2) == 2.0
    L3 = O.OracleLagrange(25.0, 0.001, 0.035)
    L3.update_lagrange_multiplier(float("nan"))
    assert np.isnan(L3.lagrangian_multiplier)


def test_dataloader_order(golden):
    g = golden("dataloader")["dataloader"]
    torch.manual_seed(g["seed"])
    for want in g["orders"]:
        assert torch.equal(O.dataloader_perm(g["S"]), want)


@pytest.mark.parametrize("kind", ["ppo", "focops"])
def test_update_chain_bit_exact(golden, kind):
    c = golden("update")["update_chain"][kind]
    pol = _policy_from(c["init"], c["D"], c["A"])
    opt = O.OracleOptim(pol)
    data, lam = c["data"], c["lam"]
    adv = (data["adv_r"] - lam * data["adv_c"]) / (lam + 1)
    res = O.pg_update(pol, opt, data, adv, kind=kind, batch_size=c["batch"], learning_iters=len(c["perms"]),
                      target_kl=1e9 if kind == "ppo" else 0.02, perms=c["perms"])
    # target_kl above only gates early stop for ppo; focops needs the real 0.02 in its indicator
    losses = torch.tensor(res["losses"])
    n = losses.shape[0]
    assert torch.equal(losses, c["losses"][:n])
    if n == c["losses"].shape[0]:
        for net in O.NET_ORDER:
            for k, v in c["final"][net].items():
                assert torch.equal(pol.nets[net][k].detach(), v), (net, k)


def test_trust_region_pieces(golden):
    c = golden("trust")["trust"]
    pol = _policy_from(c["state"], c["D"], c["A"])
    d = c["data"]
    assert list(pol.nets["actor"].keys()) == c["names"]  # log_std first
    assert torch.equal(O.flat_params(pol), c["theta"])
    Fv = O.fvp_autograd(pol, d["obs"], c["v"])
    assert torch.equal(Fv, c["Fv"])
    Fa = O.fvp_analytic(pol, d["obs"], c["v"])
    assert (Fa - c["Fv"]).norm() / c["Fv"].norm() < 1e-6
    x = O.conjugate_gradients(lambda v: O.fvp_autograd(pol, d["obs"], v), c["rhs"], 15)
    assert torch.equal(x, c["cg_x"])
    for p in pol.params("actor"):
        p.grad = None
    loss = O.surrogate_loss(pol, d["obs"], d["act"], d["log_prob"], d["adv_r"])
    loss.backward()
    assert torch.equal(loss.detach(), c["surr"]) and torch.equal(O.flat_grads(pol), c["surr_grad"])


def _load_env_mod():
    from safepo.common import synthetic_env
    return synthetic_env


@pytest.mark.parametrize("algo", ["ppo_lag", "focops", "cpo", "trpo_lag"])
def test_full_main_matches_reference(golden, algo):
    """oracle.trainers.train == the reference's main() on the same synthetic env: every
    logged number of progress.csv and the saved actor weights, bit for bit."""
    run = golden("main_runs")["main_runs"][algo]
    senv = _load_env_mod()
    args = TR.default_args(**run["args"])
    D, A = senv.TASK_DIMS[args.task]
    env = senv.SyntheticVecEnv(args.num_envs, D, A, seed=args.seed, **run["env"])
    pol, log, _ = TR.train(algo, args, env)
    assert len(log.rows) == len(run["rows"])
    for got, want in zip(log.rows, run["rows"]):
        for k, v in want.items():
            if k.startswith("Time/"):
                continue
            w = float(v.replace("tensor(", "").rstrip(")")) if isinstance(v, str) else float(v)
            g = float(got[k])
            # cpo logs Train/KL as a fp32 tensor (cpo.py:529) -> the csv holds its 8-digit repr
            assert g == w or np.float32(g) == np.float32(w) or (np.isnan(g) and np.isnan(w)), (algo, k, g, w)
    # the checkpoint the reference wrote is from epoch 0 (logger.torch_save at epoch==0)
    assert set(run["actor"].keys()) == set(pol.nets["actor"].keys())


@pytest.mark.parametrize("algo", ["ppo", "pg", "cppo_pid", "trpo", "natural_pg", "rcpo", "pcpo", "cup"])
def test_sibling_main_matches_reference(golden, algo):
    """the sibling scripts (ppo_lag / trpo_lag minus Lagrange, minus the clip, with the PID
    multiplier) -- oracle.trainers.train against the reference's own main(), every logged number bit for bit."""
    run = golden("siblings")["main_runs"][algo]
    senv = _load_env_mod()
    args = TR.default_args(**run["args"])
    D, A = senv.TASK_DIMS[args.task]
    env = senv.SyntheticVecEnv(args.num_envs, D, A, seed=args.seed, **run["env"])
    pol, log, _ = TR.train(algo, args, env)
    assert len(log.rows) == len(run["rows"])
    for got, want in zip(log.rows, run["rows"]):
        assert set(k for k in want if not k.startswith("Time/")) <= set(got.keys()), (algo, sorted(set(want) - set(got)))
        for k, v in want.items():
            if k.startswith("Time/"):
                continue
            w = float(v.replace("tensor(", "").rstrip(")")) if isinstance(v, str) else float(v)
            g = float(got[k])
            assert g == w or np.float32(g) == np.float32(w) or (np.isnan(g) and np.isnan(w)), (algo, k, g, w)
        assert ("Train/LagragianMultiplier" in got) == ("Train/LagragianMultiplier" in want), algo


def test_pid_lagrange_matches_reference(golden):
    c = golden("siblings")["pid"]
    P = O.OraclePIDLagrange(25.0, 0.001)
    for jc, want in zip(c["jc"], c["lam"]):
        P.update_lagrange_multiplier(jc)
        assert P.lagrangian_multiplier == want, (jc, P.lagrangian_multiplier, want)


def test_masked_gae_with_popart_matches_reference(golden):
    """SURVEY 8 row G2 (MAPPO-Lag, config 5): oracle/ma_oracle.py against the reference's SeparatedReplayBuffer.compute_returns /
    compute_cost_returns and PopArt, bit for bit: normalised outputs, running statistics, de-normalisation, both return tensors."""
    from oracle import ma_oracle as MA
    for c in golden("ma_gae")["ma_gae"]:
        pop = MA.OraclePopArt(1)
        for x, want in zip(c["popart_inputs"], c["popart_outputs"]):
            assert torch.equal(pop.normalize(x), want)
        for got, want in zip((pop.running_mean, pop.running_mean_sq, pop.debiasing_term), c["popart_state"]):
            assert torch.equal(got, want)
        assert torch.equal(pop.denormalize(c["value_preds"][:2]), c["denorm_probe"])
        ret = MA.masked_gae(c["rewards"], c["value_preds"], c["masks"], pop, c["gamma"], c["lam"])
        cret = MA.masked_gae(c["costs"], c["cost_preds"], c["masks"], pop, c["gamma"], c["lam"])
        assert torch.equal(ret, c["returns"][:-1]) and torch.equal(cret, c["cost_returns"][:-1]), (c["T"], c["N"])
        assert float(c["returns"][-1].abs().max()) == 0.0      # the reference's extra row stays zero


def test_ma_networks_and_ppo_update_match_reference(golden):
    """SURVEY 8f rank 3 (config 5, first slice): the oracle's MAPPO-Lag networks and ppo_update against the reference's
    own MAPPO_L_Policy / MAPPO_L_Trainer (mappolag.py:46-199): get_actions (deterministic and sampled under the same seed)
    and two consecutive updates -- every returned scalar, lamda_lagr, the PopArt state and all weights, bit for bit."""
    from oracle import ma_oracle as MA
    c = golden("ma_update")["ma_update"]
    cfg, s = c["cfg"], c["sample"]
    threads = torch.get_num_threads()
    torch.set_num_threads(1)        # like the fixture: LayerNorm's backward reductions depend on the intra-op thread count in the last bit
    nets = {n: MA.OracleMANet(c["init"][n], layer_N=cfg["layer_N"]) for n in ("actor", "critic", "cost_critic")}
    v, a, lp, k = MA.ma_get_actions(nets["actor"], nets["critic"], nets["cost_critic"], s["share_obs"], s["obs"], deterministic=True)
    d = c["actions"]["det"]
    assert torch.equal(v, d["values"]) and torch.equal(a, d["actions"]) and torch.equal(lp, d["logp"]) and torch.equal(k, d["cost"])
    torch.manual_seed(c["actions"]["sampled"]["seed"])
    v, a, lp, k = MA.ma_get_actions(nets["actor"], nets["critic"], nets["cost_critic"], s["share_obs"], s["obs"])
    d = c["actions"]["sampled"]
    assert torch.equal(a, d["actions"]) and torch.equal(lp, d["logp"]) and lp.shape == (c["dims"][3], c["dims"][2])   # per-dimension log-probs
    tr = MA.OracleMATrainer(nets["actor"], nets["critic"], nets["cost_critic"], cfg)
    for step in c["steps"]:
        out = tr.ppo_update(s)
        for name, want in step["out"].items():
            assert torch.equal(torch.as_tensor(out[name]), want), (name, out[name], want)
        assert torch.equal(torch.as_tensor(tr.lamda_lagr), step["lamda_lagr"])
        for got, want in zip((tr.popart.running_mean, tr.popart.running_mean_sq, tr.popart.debiasing_term), step["popart"]):
            assert torch.equal(got, want)
        for n in ("actor", "critic", "cost_critic"):
            for key, want in step["state"][n].items():
                assert torch.equal(nets[n].p[key].detach(), want), (n, key)
    torch.set_num_threads(threads)


def test_ma_buffer_trainer_and_runner_iteration_match_reference(golden):
    """SURVEY 8f rank 3, the callers of ppo_update: the oracle's SeparatedReplayBuffer / MAPPO_L_Trainer.train / Runner.collect,
    insert, compute, train (buffer.py:209-465, mappolag.py:200-234,402-504,583-597) against two whole iterations of two agents
    run through the reference's own methods (tests/golden/ma_epoch.pt): sampled actions and log-probs of every step (same draws
    from the global generator), masks, returns and cost returns after compute(), and after train() the cross-agent factor,
    lamda_lagr, the PopArt statistics and all weights of both agents -- bit for bit."""
    from oracle import ma_oracle as MA
    c = golden("ma_epoch")["ma_epoch"]
    T, N, D, DS, A, H, NA = c["dims"]
    cfg = c["cfg"]
    threads = torch.get_num_threads()
    torch.set_num_threads(1)
    try:
        nets = [tuple(MA.OracleMANet(st[n], layer_N=cfg["layer_N"]) for n in ("actor", "critic", "cost_critic")) for st in c["init"]]
        run = MA.OracleMARunner(nets, cfg, T, N, D, DS, A)
        run.warmup(c["obs0"], c["share_obs0"])
        torch.manual_seed(c["seed"])
        for it in range(2):
            for step in range(T):
                s = c["stream"][it][step]
                values, actions, logps, cost_preds = run.collect(step)
                assert torch.equal(values, s["values"]) and torch.equal(cost_preds, s["cost_preds"]), (it, step)
                for a in range(NA):
                    assert torch.equal(actions[a], s["actions"][a]) and torch.equal(logps[a], s["action_log_probs"][a]), (it, step, a)
                run.insert(s["obs"], s["share_obs"], s["rewards"], s["costs"], s["dones"], values, actions, logps, cost_preds)
            run.compute()
            for a in range(NA):
                want, b = c["iters"][it]["after_compute"][a], run.buffer[a]
                for k in ("returns", "cost_returns", "value_preds", "cost_preds", "masks", "active_masks"):
                    assert torch.equal(getattr(b, k), want[k]), (it, a, k)
            run.train()
            for a in range(NA):
                want = c["iters"][it]["agents"][a]
                assert torch.equal(run.buffer[a].factor, want["factor"]), (it, a)
                assert float(run.trainer[a].lamda_lagr) == float(want["lamda_lagr"]), (it, a)
                pop = run.trainer[a].popart
                for got, w in zip((pop.running_mean, pop.running_mean_sq, pop.debiasing_term), want["popart"]):
                    assert torch.equal(got.reshape(-1), w.reshape(-1)), (it, a)
                for n, net in zip(("actor", "critic", "cost_critic"), run.nets[a]):
                    for k, v in want["state"][n].items():
                        assert torch.equal(net.p[k].detach(), v), (it, a, n, k)
            if it == 0:                       # Runner.return_aver_cost after an iteration with finished episodes
                for b in run.buffer:
                    b.aver_episode_costs = torch.tensor(c["aver_cost_after_first"])
        assert any(bool((c["iters"][1]["agents"][a]["factor"] != 1).any()) for a in range(NA))     # the factor did move
    finally:
        torch.set_num_threads(threads)
