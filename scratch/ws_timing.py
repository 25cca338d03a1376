"""phase clocks of stream_kernel_ws (library built with -DGAITK_WS_TIMING, GAITK_LIB=scratch/libgaitk_timing.so)"""
import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'oracle')
import numpy as np, torch, gaitk
m = gaitk.WearGaitThreeModal().cuda(); plan = m.set_window(64).plan()
crit=[gaitk.GCLLoss(cls_num_list=[40,60],m=0.2,s=25,noise_mul=0.0) for _ in range(3)]
B=32768
rng=np.random.default_rng(0)
xs=[torch.from_numpy(rng.standard_normal((B,64,d),dtype=np.float32)).cuda() for d in (2,13,24)]
y=torch.from_numpy((rng.random(B)<0.6).astype(np.int64)).cuda()
st=gaitk.FusedTrainStep(m,crit,cagrad_c=0.5,private_mult=2.0,dtype=gaitk.DTYPE_BF16X3,process_group=False)
for i in range(2):
    st.step(xs,[y]*3); torch.cuda.synchronize(); print("---- step", i, flush=True)
