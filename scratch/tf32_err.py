import sys, numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/oracle"); sys.path.insert(0, "/root/repo/tests")
import gaitk as gk, gait_oracle as O
def relerr(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
torch.manual_seed(5)
m = gk.WearGaitThreeModal().cuda()
for B in (8, 64, 512, 4096, 32768):
    xs, y = O.synth_weargait_batch(B, seed=11)
    xs = [torch.from_numpy(x).cuda() for x in xs]; y = torch.from_numpy(y).cuda()
    for wm in ("gcl", "ce"):
        crit = [gk.GCLLoss(cls_num_list=[400, 600], m=0.2, s=25, noise_mul=0.0) if wm == "gcl" else gk.CrossEntropyLoss() for _ in range(3)]
        plan = m.set_window(64).plan(); res = {}
        for dt in (gk.DTYPE_F32, gk.DTYPE_TF32):
            st = gk.FusedTrainStep(m, crit, cagrad_c=0.5, private_mult=2.0, dtype=dt, process_group=False)
            gout = torch.zeros(plan.NP, device="cuda")
            lg = [torch.zeros(B, 2, device="cuda") for _ in range(3)]
            loss, correct = st.step(xs, [y, y, y], grads_out=gout, update=False, logits_out=lg)
            res[dt] = (torch.stack(lg).cpu().numpy(), loss.cpu().numpy(), st._gbuf[:3 * plan.P].cpu().numpy().copy(), gout.cpu().numpy(), correct.cpu().numpy())
        a, b = res[gk.DTYPE_F32], res[gk.DTYPE_TF32]
        priv = np.ones(plan.NP, bool)
        for p in plan.params:
            if p.group <= 0: priv[p.offset:p.offset+p.numel] = False
        print(f"B={B:6d} {wm:3s} logits {relerr(b[0],a[0]):.2e} loss {relerr(b[1],a[1]):.2e} G {relerr(b[2],a[2]):.2e} private {relerr(b[3][priv],a[3][priv]):.2e} "
              f"shared-final {relerr(b[3][~priv],a[3][~priv]):.2e} correct diff {np.abs(a[4]-b[4]).max():.0f}", flush=True)
