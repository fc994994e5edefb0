#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by RUNNING THE UNMODIFIED REFERENCE.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

What is executed is the reference's own code, imported from where it lies:

  * ``/root/reference/src/env/hedging_env_v2.py`` and ``hedging_env.py`` (class ``HedgingEnv``),
    through the ~60-line gymnasium stand-in in ``oracle/_gym_stub`` (gymnasium is not installed),
  * ``/root/reference/src/sim/option_price_assignment.py`` (``black_scholes_vectorized``,
    ``calculate_annualized_vol_matrix``),
  * ``/root/reference/src/tools/bs_delta.py`` (``bs_delta_hedge``),

on inputs derived from the reference's shipped ``data/paths.npy`` /
``data/paths_options.npz``.  Outputs (all small, committed):

  env_<case>.npz         per-step observation / reward / terminated / info of N reference envs
                         plus the env-schema arrays, the actions and the episode indices they used
  schema_b_golden.npz    48 shipped paths + the shipped calls/puts for them (reference known-answer pair)
  bs_delta_golden.npz    bs_delta_hedge P&L of 6 shipped paths
"""
import importlib.util
import os
import sys
import tempfile
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, os.path.join(ROOT, "oracle", "_gym_stub"))
sys.path.insert(0, ROOT)

from oracle import bs_oracle  # noqa: E402  (inputs only: ATM book columns fed to the reference env)
from oracle.hedge_oracle import INFO_FLOAT_KEYS, INFO_INT_KEYS  # noqa: E402


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def env_schema_from_paths(paths, variance_mode="realised", scale=1.0):
    """Builder's choice of the env-schema columns the reference does not ship."""
    paths = np.asarray(paths, np.float64) * scale
    if variance_mode == "realised":
        sig = bs_oracle.realised_vol_matrix(paths)
        sig[:, 0] = sig[:, 2]
        sig[:, 1] = sig[:, 2]
        var = sig ** 2
    else:
        var = np.full_like(paths, 0.02903)          # xi from estimate_base_params on the shipped CSV
    calls, puts = bs_oracle.atm_book(paths, var)
    return dict(paths=paths, volatilities=var, call_prices_atm=calls, put_prices_atm=puts)


def run_reference_envs(env_cls, data, kwargs, actions, seeds):
    """Step ``len(seeds)`` unmodified reference envs through ``actions`` (steps, n, 2), resetting on termination."""
    n = len(seeds)
    steps = actions.shape[0]
    with tempfile.TemporaryDirectory() as d:
        f = os.path.join(d, "schema_a.npz")
        np.savez(f, **data)
        envs = [env_cls(f, **kwargs) for _ in range(n)]
    obs = np.zeros((steps, n, 13), np.float32)
    reward = np.zeros((steps, n), np.float64)
    term = np.zeros((steps, n), bool)
    info_f = np.zeros((steps, n, len(INFO_FLOAT_KEYS)), np.float64)
    info_i = np.zeros((steps, n, len(INFO_INT_KEYS)), np.int64)
    reset_obs, episode_idx = [[] for _ in range(n)], [[] for _ in range(n)]
    for i, e in enumerate(envs):
        o, _ = e.reset(seed=int(seeds[i]))
        reset_obs[i].append(o)
        episode_idx[i].append(int(e.current_episode_idx))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for t in range(steps):
            for i, e in enumerate(envs):
                o, r, te, tr, inf = e.step(actions[t, i])
                assert tr is False
                obs[t, i], reward[t, i], term[t, i] = o, r, te
                info_f[t, i] = [float(inf.get(k, np.nan)) for k in INFO_FLOAT_KEYS]   # v1 lacks 3 keys -> NaN
                info_i[t, i] = [int(inf[k]) for k in INFO_INT_KEYS]
                if te:
                    o, _ = e.reset()
                    reset_obs[i].append(o)
                    episode_idx[i].append(int(e.current_episode_idx))
    n_eps = min(len(x) for x in episode_idx)
    return dict(obs=obs, reward=reward, terminated=term, info_f=info_f, info_i=info_i,
                reset_obs=np.array([[reset_obs[i][k] for i in range(n)] for k in range(n_eps)], np.float32),
                episode_idx=np.array([[episode_idx[i][k] for i in range(n)] for k in range(n_eps)], np.int64))


def main():
    env_v2 = _load("ref_env_v2", f"{REF}/src/env/hedging_env_v2.py").HedgingEnv
    env_v1 = _load("ref_env_v1", f"{REF}/src/env/hedging_env.py").HedgingEnv
    opa = _load("ref_opa", f"{REF}/src/sim/option_price_assignment.py")
    bsd = _load("ref_bsd", f"{REF}/src/tools/bs_delta.py")

    shipped = np.load(f"{REF}/data/paths.npy")
    book = np.load(f"{REF}/data/paths_options.npz")
    rng = np.random.default_rng(20261018)

    # ---------------------------------------------------------------- env cases
    n_env = 4
    seeds = 12345 + np.arange(n_env)                 # train_ppo_v2.py:78 seed base

    def uniform_actions(steps):
        return rng.uniform(-1, 1, (steps, n_env, 2)).astype(np.float32)

    def saturating_actions(steps):
        a = rng.choice(np.array([-1.0, 1.0, 1.0, 1.0, 0.3, -0.1, 0.5, 0.7, 3.0, -2.5, 0.0], np.float32),
                       size=(steps, n_env, 2)).astype(np.float32)
        a[5, 0, 0] = np.nan
        a[6, 1, 1] = np.inf
        a[7, 2, 0] = -np.inf
        a[8, 3, 1] = 1e30
        return a

    full = env_schema_from_paths(shipped[:24])
    short = {k: v[:, : (41 if v.shape[1] == 253 else 40)] for k, v in env_schema_from_paths(shipped[24:40]).items()}
    const_var = {k: v[:, : (41 if v.shape[1] == 253 else 40)]
                 for k, v in env_schema_from_paths(shipped[40:56], "const").items()}
    low = {k: v[:, : (41 if v.shape[1] == 253 else 40)]
           for k, v in env_schema_from_paths(shipped[56:72], "realised", scale=0.04).items()}
    tiny = {k: v.copy() for k, v in low.items()}
    tiny["paths"][0, :] = 0.0                         # S <= 1e-6 branch of _calculate_greeks, S0 -> 1.0
    tiny["paths"][1, 3:9] = 1e-7
    tiny["volatilities"][2, :] = 0.0                  # sigma floor sqrt(1e-8)
    tiny["paths"][3, :] = 0.4                         # K = round(S) = 0 -> K_checked = 1e-6

    cases = {
        # the training configuration of train_ppo_v2.py:74-80 on full-length episodes
        "v2_train": (env_v2, full, dict(slippage_bps=1.0, theta_weight=2e-4, pnl_penalty_weight=1e-3,
                                        lambda_cost=1e-4, loss_type="abs"), uniform_actions(2 * 252 + 9)),
        # v1 defaults (hedging_env.py:10-20)
        "v1_default": (env_v1, short, dict(), uniform_actions(3 * 40 + 7)),
        # mse loss, non-zero initial cash, clamps at +-200 / NaN / inf / out-of-range actions
        "v2_mse_saturating": (env_v2, short, dict(loss_type="mse", pnl_penalty_weight=0.5, initial_cash=1000.0,
                                                  slippage_bps=2.5, theta_weight=1e-3), saturating_actions(3 * 40 + 7)),
        # cvar string falls back to abs; no greeks
        "v2_cvar_nometrics": (env_v2, const_var, dict(loss_type="cvar", record_metrics=False,
                                                      shares_to_hedge=5000, max_contracts_held_per_type=40,
                                                      max_trade_per_step=7), uniform_actions(2 * 40 + 5)),
        # S0 below the 25 floor
        "v2_lowprice": (env_v2, low, dict(slippage_bps=1.0, transaction_cost_per_contract=0.05),
                        uniform_actions(2 * 40 + 5)),
        # degenerate prices / variances
        "v2_degenerate": (env_v2, {k: v[:4] for k, v in tiny.items()}, dict(slippage_bps=1.0),
                          uniform_actions(2 * 40 + 5)),
    }
    for name, (cls, data, kwargs, actions) in cases.items():
        out = run_reference_envs(cls, data, kwargs, actions, seeds)
        kw_items = {f"kw_{k}": np.array(v) for k, v in kwargs.items()}
        np.savez_compressed(os.path.join(HERE, f"env_{name}.npz"), version=np.array(1 if cls is env_v1 else 2),
                            actions=actions, seeds=seeds, **data, **out, **kw_items)
        print(f"env_{name}: steps={actions.shape[0]} episodes={out['episode_idx'].shape[0]} "
              f"terminations={int(out['terminated'].sum())}")

    # ------------------------------------------------- schema B known-answer pair
    sl = slice(0, 48)
    vols = opa.calculate_annualized_vol_matrix(shipped[sl])
    strikes = np.round(shipped[sl, 0])
    Tgrid = np.clip(1 - np.arange(shipped.shape[1]) / 252, 0, None)
    calls = np.zeros_like(shipped[sl])
    puts = np.zeros_like(shipped[sl])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for t in range(shipped.shape[1]):
            calls[:, t], puts[:, t] = opa.black_scholes_vectorized(shipped[sl, t], strikes, Tgrid[t], 0.04, vols[:, t])
    err = np.nanmax(np.abs(calls - book["calls"][sl]))
    assert err < 1e-11 and np.array_equal(np.isnan(calls), np.isnan(book["calls"][sl])), err
    np.savez_compressed(os.path.join(HERE, "schema_b_golden.npz"), paths=shipped[sl],
                        calls_shipped=book["calls"][sl], puts_shipped=book["puts"][sl],
                        calls_recomputed=calls, puts_recomputed=puts, vols=vols)
    print(f"schema_b_golden: reference functions reproduce the shipped npz to {err:.2e} abs")

    # ------------------------------------------------------------ bs_delta golden
    pnl = bsd.bs_delta_hedge(shipped[:6])
    np.savez_compressed(os.path.join(HERE, "bs_delta_golden.npz"), paths=shipped[:6], pnl=pnl)
    print("bs_delta_golden: 6 paths")


def make_outer_euler_golden():
    """Run the unmodified generate_paths_and_options (rbergomi_sim.py:309-499) on CPU under the cupy stand-in, with the
    module constants shrunk, and export its per-path parameters, variances, unscaled draws and resulting paths."""
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_cupy_stub"))
    rb = _load("ref_rbergomi", f"{REF}/src/sim/rbergomi_sim.py")
    rb.N_STEPS, rb.N_PATHS_OPTION_MC, rb.OPTION_PRICING_MINI_BATCH_SIZE = 12, 40, 16
    rb.tqdm = lambda it, **kw: _NoBar(it)
    hist = np.loadtxt(f"{REF}/data/historical_prices.csv", dtype=np.float64, delimiter=",")
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as d:
        os.chdir(d)
        try:
            import contextlib
            import io
            with contextlib.redirect_stdout(io.StringIO()):
                paths, v, calls, puts = rb.generate_paths_and_options(hist, 24, rb.R, rb.DT, 42)
            chk = dict(np.load(rb.CHECKPOINT_FILE, allow_pickle=True))
        finally:
            os.chdir(cwd)
    np.savez_compressed(os.path.join(HERE, "outer_euler_golden.npz"), paths=np.asarray(paths), v=np.asarray(v),
                        S0=chk["S0_arr_gpu"], rho=chk["rho_arr_gpu"], dW1=chk["dW1_unscaled_main_gpu"],
                        dW2=chk["dW2_unscaled_main_gpu"], r=rb.R, dt=rb.DT)
    print(f"outer_euler_golden: {paths.shape[0]} paths x {paths.shape[1] - 1} days from the unmodified simulator")


def make_policy_golden():
    """Run the unmodified hand-written policies (baselines.py:74-103, delta_and_nothing.py:122-163) on random observations."""
    import types
    bl = _load("ref_baselines", f"{REF}/src/agents/baselines.py")
    dn = _load("ref_delta_and_nothing", f"{REF}/src/benchmark/delta_and_nothing.py")
    rng = np.random.default_rng(20240)
    n = 600
    obs = rng.uniform(-1, 1, (n, 13)).astype(np.float32)
    obs[:, 7] = rng.uniform(0, 1, n)
    obs[:, 9] = obs[:, 7] - 1
    obs[::7, 7] = 0.0
    obs[::11, 9] = 0.0
    obs[::13, 7] = 5e-4                                    # |delta * 100| below the 0.1 threshold
    pos_c, pos_p = rng.integers(-200, 201, n), rng.integers(-200, 201, n)
    pos_c[:60] = 0
    pos_p[:60] = 0
    obs[:, 3], obs[:, 4] = pos_c / 200.0, pos_p / 200.0
    env = types.SimpleNamespace(max_contracts_held=200, option_contract_multiplier=100, shares_held_fixed=10000,
                                max_trade_per_step=15, action_space=types.SimpleNamespace(dtype=np.float32))
    a_every = np.stack([bl.policy_delta_every_step(o, env) for o in obs])
    a_none = np.stack([bl.policy_no_hedge(o, env) for o in obs])
    a_bench = np.stack([dn.delta_hedging_action_selector(dict(
        S_t=100.0, v_t=0.04, call_delta_atm=o[7], put_delta_atm=o[9], current_call_contracts=np.int64(c),
        current_put_contracts=np.int64(q), shares_to_hedge=10000, option_contract_multiplier=100, max_trade_per_step=15))
        for o, c, q in zip(obs, pos_c, pos_p)])
    np.savez_compressed(os.path.join(HERE, "policy_golden.npz"), obs=obs, pos_c=pos_c, pos_p=pos_p,
                        delta_every_step=a_every, no_hedge=a_none, delta_benchmark=a_bench)
    print(f"policy_golden: {n} observations through the unmodified policies")


def make_rbergomi_golden():
    """Run the unmodified rBergomi building blocks (rbergomi_sim.py:206-306, 357-406, 454-464) on the CPU under the cupy
    stand-in and export inputs, draws and outputs: the nested-MC pricer on a small batch, and the full generator."""
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_cupy_stub"))
    rb = _load("ref_rbergomi2", f"{REF}/src/sim/rbergomi_sim.py")
    import cupy as cp
    rng = np.random.default_rng(77)
    B, n_mc = 6, 96
    S0 = rng.uniform(80, 520, B)
    K = np.round(S0)
    xi = rng.uniform(0.01, 0.09, B)
    H = np.array([0.01, 0.07, 0.1, 0.25, 0.4656, 0.49])
    eta = rng.uniform(0.5, 2.2, B)
    rho = rng.uniform(-0.95, -0.05, B)
    out = dict(S0=S0, K=K, xi=xi, H=H, eta=eta, rho=rho, r=rb.R, dt=rb.DT, tenor=rb.T_OPTION_TENOR)
    for kind in ("call", "put"):
        cp.random.seed(1000 + len(kind))
        price = rb.price_rbergomi_option_gpu(S0, K, rb.T_OPTION_TENOR, rb.R, xi, H, eta, rho, kind, n_mc, rb.DT)
        cp.random.seed(1000 + len(kind))                       # the same draws, in the order the function takes them (:271-272)
        M = rb.next_power_of_two(int(rb.T_OPTION_TENOR / rb.DT) + 1)
        Z = cp.random.normal(size=(B, n_mc, M)) + 1j * cp.random.normal(size=(B, n_mc, M))
        out[f"Z_{kind}"], out[f"price_{kind}"] = Z, np.asarray(price)
    # the generator itself, constants shrunk (12 days, 40 inner paths): per-path parameters, main increments, variance, paths
    rb.N_STEPS, rb.N_PATHS_OPTION_MC, rb.OPTION_PRICING_MINI_BATCH_SIZE = 12, 40, 16
    rb.tqdm = lambda it, **kw: _NoBar(it)
    hist = np.loadtxt(f"{REF}/data/historical_prices.csv", dtype=np.float64, delimiter=",")
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as d:
        os.chdir(d)
        try:
            import contextlib
            import io
            with contextlib.redirect_stdout(io.StringIO()):
                paths, v, calls, puts = rb.generate_paths_and_options(hist, 24, rb.R, rb.DT, 42)
            chk = dict(np.load(rb.CHECKPOINT_FILE, allow_pickle=True))
        finally:
            os.chdir(cwd)
    out.update(gen_paths=np.asarray(paths), gen_v=np.asarray(v), gen_S0=chk["S0_arr_gpu"], gen_xi=chk["xi_arr_gpu"], gen_H=chk["H_arr_gpu"],
               gen_eta=chk["eta_arr_gpu"], gen_rho=chk["rho_arr_gpu"], gen_dW1=chk["dW1_unscaled_main_gpu"],
               gen_dW2=chk["dW2_unscaled_main_gpu"], gen_Z=chk["Z_main_gpu"],
               gen_base=np.array([chk["S0_base"], chk["xi_base"], chk["H_base"], chk["eta_base"], chk["rho_base"]], np.float64),
               gen_n_steps=12)
    np.savez_compressed(os.path.join(HERE, "rbergomi_golden.npz"), **out)
    print(f"rbergomi_golden: pricer batch of {B} x {n_mc} inner paths (call, put) + generator 24 paths x 12 days")


def make_lstm_golden():
    """The shipped recurrent policy run by the reference's OWN network class on observations of the reference env.

    ``quantconnect/model_wrapper.py: RecurrentPPOModel`` (LSTM(13, 128) -> Linear(128, 64) ReLU Linear(64, 64) ReLU ->
    Linear(64, 2) -> tanh) with ``quantconnect/model_files/policy_weights.pth`` loaded through the key mapping of
    ``ModelWrapper.LoadModel`` (:84-103) and the normalisation of ``ModelWrapper.predict`` (:131) with
    ``normalization_stats.pkl``.  The hidden state is carried from step to step and zeroed after a terminated step (the
    wrapper's ``reset_hidden_states``).  Inputs: steps 220-299 of the 4 reference envs of env_v2_train.npz (an episode boundary
    inside), brought inside +-9.5 sigma of the shipped statistics.  The weights travel in the fixture: /root/reference does not exist on the GPU box."""
    import pickle
    import types

    import torch
    sys.modules.setdefault("AlgorithmImports", types.ModuleType("AlgorithmImports"))          # LEAN's star-import module
    mw = _load("ref_model_wrapper", os.path.join(REF, "quantconnect", "model_wrapper.py"))
    sd = torch.load(os.path.join(REF, "quantconnect", "model_files", "policy_weights.pth"), map_location="cpu", weights_only=False)
    stats = pickle.load(open(os.path.join(REF, "quantconnect", "model_files", "normalization_stats.pkl"), "rb"))
    model = mw.RecurrentPPOModel()
    ours = {"lstm_actor." + k: sd["lstm_actor." + k] for k in ("weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0")}
    for j in (0, 2):
        for kind in ("weight", "bias"):
            ours[f"mlp_extractor_policy_net.{j}.{kind}"] = sd[f"mlp_extractor.policy_net.{j}.{kind}"]
    ours["action_net.weight"], ours["action_net.bias"], ours["log_std"] = sd["action_net.weight"], sd["action_net.bias"], sd["log_std"]
    model.load_state_dict(ours)
    model.eval()
    env = np.load(os.path.join(HERE, "env_v2_train.npz"))
    obs = np.concatenate([env["reset_obs"][:1], env["obs"][:300]], axis=0)[220:300].astype(np.float32)   # [80, 4, 13], crosses t = 252
    done = np.concatenate([np.zeros((1, 4), bool), env["terminated"][:300].astype(bool)], axis=0)[220:300]
    mean, var = np.asarray(stats["obs_mean"], np.float64), np.asarray(stats["obs_var"], np.float64)
    # these envs run on paths.npy-derived data, not on the training set: the deltas reach 12 sigma and the variance change 35.
    # The wrapper never clips (SB3's VecNormalize did, at 10, in training), so bring the inputs inside +-9.5 sigma: then the
    # reference network, the oracle and the kernel (which clips at 10) all see the same thing.
    obs = (mean + np.sqrt(var + 1e-8) * np.clip((obs - mean) / np.sqrt(var + 1e-8), -9.5, 9.5)).astype(np.float32)
    n_steps, n = obs.shape[:2]
    hid = (torch.zeros(1, n, 128), torch.zeros(1, n, 128))
    acts = np.zeros((n_steps, n, 2), np.float32)
    worst = 0.0
    with torch.no_grad():
        for t in range(n_steps):
            x = (obs[t] - mean) / np.sqrt(var + 1e-8)                                          # model_wrapper.py:131
            worst = max(worst, float(np.abs(x).max()))
            xt = torch.as_tensor(x, dtype=torch.float32).unsqueeze(1)                          # [n, 1, 13], batch_first
            a, hid = model(xt, hid)
            acts[t] = a[:, 0].numpy()
            fin = torch.as_tensor(done[t])
            hid = (hid[0] * (~fin).float()[None, :, None], hid[1] * (~fin).float()[None, :, None])
    assert worst < 10.0, worst            # VecNormalize's clip_obs = 10 never bites on these inputs
    np.savez_compressed(os.path.join(HERE, "lstm_golden.npz"), obs=obs, done=done, actions=acts, obs_mean=mean.astype(np.float32),
                        obs_var=var.astype(np.float32), max_abs_normalised_obs=worst,
                        w_ih=sd["lstm_actor.weight_ih_l0"].numpy(), w_hh=sd["lstm_actor.weight_hh_l0"].numpy(),
                        b_ih=sd["lstm_actor.bias_ih_l0"].numpy(), b_hh=sd["lstm_actor.bias_hh_l0"].numpy(),
                        W1=sd["mlp_extractor.policy_net.0.weight"].numpy(), b1=sd["mlp_extractor.policy_net.0.bias"].numpy(),
                        W2=sd["mlp_extractor.policy_net.2.weight"].numpy(), b2=sd["mlp_extractor.policy_net.2.bias"].numpy(),
                        W3=sd["action_net.weight"].numpy(), b3=sd["action_net.bias"].numpy())
    print(f"lstm_golden: {n_steps} steps x {n} envs through the reference's RecurrentPPOModel with the shipped weights; "
          f"max |normalised obs| = {worst:.2f}, actions in [{acts.min():.3f}, {acts.max():.3f}]")


def make_mlp_golden():
    """Pins the plain MLP policy (BASELINE configs[4]: 13-64-64-2) on the reference's own network and shipped weights.

    The reference ships no 13-input MLP: its head ``mlp_extractor.policy_net`` (Linear(128, 64) ReLU Linear(64, 64) ReLU) +
    ``action_net`` (Linear(64, 2)) + tanh (quantconnect/model_wrapper.py:177-185, 202) reads the 128 LSTM outputs.  Following
    SURVEY section 8(d) (C5), the head is fed from a FIXED 13 -> 128 linear projection P of the normalised observation
    (seeded, stored in the fixture): actions = tanh(action_net(policy_net(P x))), evaluated by the reference's own
    ``RecurrentPPOModel`` sub-modules with ``policy_weights.pth`` loaded.  Because P and the first head layer are both linear,
    this network IS a 13-64-64-2 ReLU MLP with first layer W1 P -- which is what the kernels run.  Observations, normalisation
    statistics and head weights are those of lstm_golden.npz (same generator conventions)."""
    import types

    import torch
    sys.modules.setdefault("AlgorithmImports", types.ModuleType("AlgorithmImports"))
    mw = _load("ref_model_wrapper", os.path.join(REF, "quantconnect", "model_wrapper.py"))
    sd = torch.load(os.path.join(REF, "quantconnect", "model_files", "policy_weights.pth"), map_location="cpu", weights_only=False)
    model = mw.RecurrentPPOModel()
    ours = {"lstm_actor." + k: sd["lstm_actor." + k] for k in ("weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0")}
    for j in (0, 2):
        for kind in ("weight", "bias"):
            ours[f"mlp_extractor_policy_net.{j}.{kind}"] = sd[f"mlp_extractor.policy_net.{j}.{kind}"]
    ours["action_net.weight"], ours["action_net.bias"], ours["log_std"] = sd["action_net.weight"], sd["action_net.bias"], sd["log_std"]
    model.load_state_dict(ours)
    model.eval()
    g = np.load(os.path.join(HERE, "lstm_golden.npz"))
    obs = g["obs"].reshape(-1, 13)                                                             # 320 observations of the reference env
    mean, var = g["obs_mean"].astype(np.float64), g["obs_var"].astype(np.float64)
    P = (0.3 * np.random.default_rng(128).standard_normal((128, 13)) / np.sqrt(13)).astype(np.float32)   # keeps most actions off the tanh plateaus
    x = ((obs - mean) / np.sqrt(var + 1e-8)).astype(np.float32)                                # model_wrapper.py:131
    with torch.no_grad():
        feats = model.mlp_extractor_policy_net(torch.as_tensor(x @ P.T))
        means = model.action_net(feats)
        acts = torch.tanh(means).numpy()                                                       # model_wrapper.py:202
    np.savez_compressed(os.path.join(HERE, "mlp_golden.npz"), obs=obs, projection=P, action_means=means.numpy(), actions=acts)
    print(f"mlp_golden: {obs.shape[0]} observations through the reference's shipped head behind a fixed 13->128 projection; "
          f"actions in [{acts.min():.3f}, {acts.max():.3f}], std {acts.std():.3f}")


def make_calibration_golden():
    """estimate_base_params (rbergomi_sim.py:171-193) of the unmodified reference on the shipped price history and on
    synthetic histories of several lengths (short ones exercise the default / guard branches)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_cupy_stub"))
    rb = _load("ref_rbergomi3", f"{REF}/src/sim/rbergomi_sim.py")
    hist = np.loadtxt(f"{REF}/data/historical_prices.csv", dtype=np.float64, delimiter=",")
    rng = np.random.default_rng(31)
    series = {"shipped": hist}
    for n in (5, 21, 25, 60, 333, 2500):
        series[f"synthetic_{n}"] = 100 * np.exp(np.cumsum(rng.normal(0, 0.01 * (1 + 0.5 * np.sin(np.arange(n) / 7)), n)))
    series["flat_40"] = np.full(40, 50.0)
    out = {}
    for k, p in series.items():
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out[f"prices_{k}"] = p
            out[f"params_{k}"] = np.array([float(x) for x in rb.estimate_base_params(p, rb.DT)], np.float64)
    np.savez_compressed(os.path.join(HERE, "calibration_golden.npz"), **out)
    print(f"calibration_golden: {len(series)} price histories; shipped -> {out['params_shipped']}")


class _NoBar:
    def __init__(self, it):
        self.it = it

    def __enter__(self):
        return self.it

    def __exit__(self, *a):
        return False

    def __iter__(self):
        return iter(self.it)


if __name__ == "__main__":
    if "--calibration-only" in sys.argv:
        make_calibration_golden()
        sys.exit(0)
    if "--rbergomi-only" in sys.argv:
        make_rbergomi_golden()
        sys.exit(0)
    if "--lstm-only" in sys.argv:
        make_lstm_golden()
        sys.exit(0)
    if "--mlp-only" in sys.argv:
        make_mlp_golden()
        sys.exit(0)
    if "--policy-only" in sys.argv:
        make_policy_golden()
        sys.exit(0)
    if "--euler-only" not in sys.argv:
        main()
        make_policy_golden()
    make_outer_euler_golden()
    make_rbergomi_golden()
    make_calibration_golden()
    make_lstm_golden()
    make_mlp_golden()
