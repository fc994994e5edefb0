import os, sys
import numpy as np, torch
root = os.environ.get("TREE") or os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root); sys.path.insert(0, root + "/tests")
from cantorrl_b200.rollout import HedgingRollout, pack_lstm
from oracle import rollout_oracle
from test_rollout_gpu import _lstm_weights, _book, KW
n_paths, T, n_envs, n_steps = 61, 12, int(sys.argv[1]) if len(sys.argv) > 1 else 203, 41
S, V, C, P = _book(n_paths, T, heston=True)
w = _lstm_weights()
g = np.random.default_rng(1)
mean, var = g.normal(0, 0.2, 13).astype(np.float32), g.uniform(0.05, 2.0, 13).astype(np.float32)
ro = HedgingRollout(data=dict(paths=S, volatilities=V, call_prices_atm=C, put_prices_atm=P), num_envs=n_envs, **KW)
stats = ro.new_stats()
res = ro.run(n_steps, "lstm_bf16", mlp=pack_lstm(**w, obs_mean=mean, obs_var=var), stats=stats, store=True)
torch.cuda.synchronize()
got = {k: getattr(res, k).cpu().numpy() for k in ("obs", "actions", "reward", "done")}
want = rollout_oracle.lstm_actor_sequence(got["obs"], got["done"], **w, mean=mean, var=var, bf16=True)
err = np.abs(got["actions"] - want).max(axis=2)
print("flag", float(stats.sums[15]), "max", err.max(), "mean", err.mean())
print("per step max:", np.round(err.max(axis=1), 4))
print("per env-block(32) max:", np.round(err.reshape(n_steps, -1)[:, : (n_envs // 32) * 32].reshape(n_steps, -1, 32).max(axis=(0, 2)), 4))

np.save(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "lstm_actions_%s.npy" % os.environ.get("TAG", "cur")), got["actions"])
