import sys; sys.path.insert(0, "/root/repo")
import gaitk, torch
m = gaitk.WearGaitThreeModal().cuda().set_window(64); p = m.plan()
for s in range(3):
    print(s, "f32", p.geometry(s, gaitk.DTYPE_F32), "tf32", p.geometry(s, gaitk.DTYPE_TF32))
