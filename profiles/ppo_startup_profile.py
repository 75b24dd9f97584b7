"""Start-up cost of ppo.train (first call of a process, then a cProfile of the second call): 0.46 s / 0.33 s for four
ppo-sa updates on a B200, i.e. about 0.12 s of graph capture + set-up per run and 0.14 s of one-time lazy loading."""
import cProfile, os, pstats, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rsoccer_isaac_cleanrl_b200 import ppo
torch.zeros(1, device="cuda"); torch.cuda.synchronize()
args = ppo.parse_args(["--env-id", "sa", "--num-envs", "4096", "--quiet", "--total-timesteps", str(4096*128*4)])
t0=time.time(); ppo.train(args); print("first train() call", time.time()-t0)
pr = cProfile.Profile(); t0=time.time(); pr.enable(); st = ppo.train(args); pr.disable(); print("second train() call", time.time()-t0, st["rollout_wall"], st["update_wall"])
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
