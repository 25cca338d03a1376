import sys, torch
sys.path.insert(0, ".")
import gaitk as gk
def build(graph):
    torch.manual_seed(9)
    m = gk.WearGaitThreeModal(synchronized=True).cuda()
    crit = [gk.GCLLoss(cls_num_list=[40, 60], m=0.2, s=25.0, noise_mul=0.0) for _ in range(3)]
    return m, gk.FusedTrainStep(m, crit, cagrad_c=0.5, private_mult=2.0, use_graph=graph)
ma, sa = build(True); mb, sb = build(False)
g = torch.Generator(device="cuda").manual_seed(1)
for it in range(14):
    B = 40 if it % 3 else 24
    xs = [torch.rand(B, 64, 2, device="cuda", generator=g), torch.randn(B, 64, 13, device="cuda", generator=g), torch.randn(B, 64, 24, device="cuda", generator=g)]
    y = torch.randint(0, 2, (B,), device="cuda", generator=g); y[0], y[1] = 0, 1
    xc = [x.clone() for x in xs]; yc = y.clone()
    ng = len(sa._graphs); ns = len(sa.__dict__.get("_raw_seen", {}))
    la, _ = sa.step(xc, [yc] * 3); la = la.clone()
    lb, _ = sb.step(xs, [y, y, y]); lb = lb.clone()
    torch.cuda.synchronize()
    print(it, B, "graphs", ng, "->", len(sa._graphs), "seen", ns, "->", len(sa._raw_seen), "ptrs", [hex(t.data_ptr())[-7:] for t in xc], hex(yc.data_ptr())[-7:],
          "OK" if torch.equal(la, lb) else f"MISMATCH {la.tolist()} {lb.tolist()}", "params equal", torch.equal(ma.flat_params(), mb.flat_params()))
