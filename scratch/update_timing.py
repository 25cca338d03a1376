"""phase clocks of cagrad_update_kernel (timing build: -DGAITK_UPDATE_TIMING -> diag[17..19] = Gram, solve + wait, apply)"""
import sys, torch
sys.path.insert(0, ".")
import gaitk as gk
torch.manual_seed(0)
for B in (64, 4096):
    m = gk.WearGaitThreeModal(synchronized=True).cuda()
    crit = [gk.GCLLoss(cls_num_list=[40, 60], m=0.2, s=25.0, noise_mul=0.0) for _ in range(3)]
    st = gk.FusedTrainStep(m, crit, cagrad_c=0.5, private_mult=2.0, use_graph=False)
    xs = [torch.rand(B, 64, 2, device="cuda"), torch.randn(B, 64, 13, device="cuda"), torch.randn(B, 64, 24, device="cuda")]
    y = torch.randint(0, 2, (B,), device="cuda"); y[0], y[1] = 0, 1
    for it in range(6):
        st.step(xs, [y, y, y]); torch.cuda.synchronize()
        d = st.diag().cpu().numpy()
        print(f"B={B} it={it} clocks gram {d[17]:.0f} solve {d[18]:.0f} apply {d[19]:.0f} | slsqp iters {d[14]:.0f} w {d[:3]}")
