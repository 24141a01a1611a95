import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, math
from vo_single_camera_sos_b200 import synth, workload, ops
from vo_single_camera_sos_b200.frontend import FrontendConfig
from oracle import driver as odriver
from test_gpu_driver import trimmed
def build_cpu(name,batch,n_frames,seed,score_mode):
    c=workload.CONFIGS[name]
    rig=synth.make_rig(c["width"],c["height"],c["pano_cols"],seed=seed)
    scene=synth.make_scene(int(c["feat"]*2.0),seed=seed)
    p=rig.pano
    H=c["n_hyp"]
    hyp_host=np.random.default_rng(seed+7).integers(0,2**32,(H,3),dtype=np.uint64).astype(np.uint32)
    thr=1.0-math.cos(math.radians(5.0))
    cfg=FrontendConfig(batch=batch,src_h=rig.height,src_w=rig.width,pano_rows=p["rows"],pano_cols=p["cols"],pano_top=rig.pano_vector(),pano_bot=rig.pano_vector(),f_top=rig.f_top,f_bot=rig.f_bot,max_feat_per_view=c["cap"],max_feat_per_bucket=c["max_bucket"],cap=c["cap"],n_hyp=H,score_mode=score_mode,ransac_threshold=thr)
    traj=synth.make_trajectory(n_frames,seed=seed)
    return workload.Workload(name,rig,scene,cfg,None,None,hyp_host,{}, {}, traj)
for seed in (5,6,7):
  w=build_cpu("tiny",4,15,seed,ops.SCORE_BEARING)
  fr=workload.make_frames(w,0,15,render=False)
  rig=np.zeros((2,3,4)); rig[:,:,:3]=np.eye(3); rig[0,:,3]=w.rig.f_top; rig[1,:,3]=w.rig.f_bot
  for pm in (0.05,0.055,0.06,0.065,0.07):
    th=dict(odriver.INDOOR,pos_min=pm)
    want=odriver.run_vo([trimmed(fr,i) for i in range(15)],(dict(w.rig.pano),w.rig.f_top,w.rig.f_bot,w.cfg.cap),(w.hyp_host,"bearing",w.cfg.ransac_threshold,rig,0.125*0.5*w.cfg.pano_cols),thresholds=th)
    m=min(min(abs(d-pm),abs(d-0.2)) for d,a,*_ in want["decisions"])
    print(seed,pm,want["status"],want["keyframe_ids"],'margin %.4f'%m)
