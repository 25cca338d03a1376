import sys; sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/oracle")
import gaitk, torch, gait_oracle as O
m = gaitk.WearGaitThreeModal().cuda()
B = 32768
xs, y = O.synth_weargait_batch(B, seed=1)
xs = [torch.from_numpy(x).cuda() for x in xs]; y = torch.from_numpy(y).cuda()
crit = [gaitk.GCLLoss(cls_num_list=[400, 600], m=0.2, s=25, noise_mul=0.0) for _ in range(3)]
st = gaitk.FusedTrainStep(m, crit, cagrad_c=0.5, private_mult=2.0, dtype=gaitk.DTYPE_TF32, process_group=False)
for i in range(2):
    st.step(xs, [y, y, y]); torch.cuda.synchronize()
    print("---- step", i, flush=True)
