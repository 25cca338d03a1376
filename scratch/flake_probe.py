"""How close to its tolerance does test_fused_step_matches_oracle_random run?  Prints the worst parameter error of
`reps` repetitions of the test body, GPU side repeated (is the GPU path deterministic?) and CPU oracle side once."""
import sys, numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "oracle"); sys.path.insert(0, "tests")
import gaitk as gk, gait_oracle as O
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
torch.manual_seed(3)
m0 = gk.WearGaitThreeModal(synchronized=False)
state = {k: v.detach().cpu().numpy().copy() for k, v in m0.state_dict().items()}
B = 201
xs, y = O.synth_weargait_batch(B, seed=9)
r = np.random.default_rng(1); ys = [y, r.permutation(y), r.permutation(y)]
counts = [[40, 90], [55, 60], [20, 30]]
p = O.canonical_params(state, False); bufs = {}
for it in range(2):
    ex = O.weargait_train_step(p, bufs, [torch.from_numpy(x) for x in xs], [torch.from_numpy(v) for v in ys],
                               synchronized=False, wm="gcl", counts=counts, alpha=0.5)
outs = []
for rep in range(reps):
    m = gk.WearGaitThreeModal(synchronized=False)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in state.items()}); m = m.cuda()
    crit = [gk.GCLLoss(cls_num_list=c, m=0.2, s=25, noise_mul=0.0) for c in counts]
    step = gk.FusedTrainStep(m, crit, cagrad_c=0.5, private_mult=2.0)
    for it in range(2):
        step.step([dev(x) for x in xs], [dev(v) for v in ys])
    sd = {k: v.cpu().numpy().astype(np.float64) for k, v in m.state_dict().items()}
    outs.append(sd)
    worst = max(((np.abs(sd[k] - p[k].detach().numpy()).max() / max(np.abs(p[k].detach().numpy()).max(), 1e-30)), k) for k in sd if k in p)
    same = all(np.array_equal(sd[k], outs[0][k]) for k in sd)
    print(f"rep {rep}: worst rel err {worst[0]:.3e} at {worst[1]}; identical to rep 0: {same}")
    if not same:
        for k in sd:
            if not np.array_equal(sd[k], outs[0][k]): print("   differs:", k, np.abs(sd[k] - outs[0][k]).max())
