import sys, numpy as np, torch, time
sys.path.insert(0,'.'); sys.path.insert(0,'oracle'); sys.path.insert(0,'tests')
import gaitk, gait_oracle as O
def relerr(a,b):
    a=np.asarray(a,dtype=np.float64); b=np.asarray(b,dtype=np.float64); return float(np.linalg.norm(a-b)/max(np.linalg.norm(b),1e-30))
torch.manual_seed(3)
m = gaitk.WearGaitThreeModal().cuda()
plan = m.set_window(64).plan()
for s in range(3): print("geom", s, plan.geometry(s, gaitk.DTYPE_BF16X3))
crit=[gaitk.GCLLoss(cls_num_list=[40,60],m=0.2,s=25,noise_mul=0.0) for _ in range(3)]
for B in (8, 64, 1000):
    xs,y=O.synth_weargait_batch(B,seed=9)
    xd=[torch.from_numpy(x).cuda() for x in xs]; yd=torch.from_numpy(y).cuda()
    res={}
    for dt in (gaitk.DTYPE_F32, gaitk.DTYPE_TF32, gaitk.DTYPE_BF16X3):
        st=gaitk.FusedTrainStep(m,crit,cagrad_c=0.5,private_mult=2.0,dtype=dt,process_group=False)
        gout=torch.zeros(plan.NP,device='cuda')
        lg=[torch.zeros(B,2,device='cuda') for _ in range(3)]
        loss,cor=st.step(xd,[yd]*3,grads_out=gout,update=False,logits_out=lg)
        torch.cuda.synchronize()
        res[dt]=(torch.stack(lg).cpu().numpy(),loss.cpu().numpy(),st._gbuf[:3*plan.P].cpu().numpy(),gout.cpu().numpy(),cor.cpu().numpy())
    a=res[gaitk.DTYPE_F32]
    for nm,dt in (("tf32",gaitk.DTYPE_TF32),("bf16x3",gaitk.DTYPE_BF16X3)):
        b=res[dt]
        print(B,nm,"logits %.2e loss %.2e G %.2e grads %.2e"%(relerr(b[0],a[0]),relerr(b[1],a[1]),relerr(b[2],a[2]),relerr(b[3],a[3])),"correct",b[4],a[4], flush=True)
    if B==8:
        got={p.name:res[gaitk.DTYPE_BF16X3][3][p.offset:p.offset+p.numel] for p in plan.params}
        ref={p.name:a[3][p.offset:p.offset+p.numel] for p in plan.params}
        for k in got: print("   ",k,"%.2e"%relerr(got[k],ref[k]) if np.linalg.norm(ref[k])>0 else "zero")
