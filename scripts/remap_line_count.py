"""Count the distinct 128-byte lines one warp-wide tap load touches for several lane-to-pixel mappings, on the C2 LUT
(CPU only; backs the remap analysis in DESIGN.md section 5)."""
import sys; sys.path.insert(0,'/root/repo')
import numpy as np
from vo_single_camera_sos_b200 import synth
from vo_single_camera_sos_b200.workload import CONFIGS
from oracle import geometry as G
c=CONFIGS['c2']
rig=synth.make_rig(c["width"],c["height"],c["pano_cols"],seed=0)
p=rig.pano; rows,cols=p["rows"],p["cols"]
print(rows,cols,rig.width,rig.height)
for which in ("top","bot"):
    lo,hi = rig.elev_top if which=="top" else rig.elev_bot
    g = rig.gum_top if which=="top" else rig.gum_bot
    mx,my=G.lut_build(g,rows,cols,p["cyl_height_max"],p["cyl_height_min"],lo,hi)
    ok=np.isfinite(mx)&np.isfinite(my)
    x0=np.floor(np.nan_to_num(mx)).astype(np.int64); y0=np.floor(np.nan_to_num(my)).astype(np.int64)
    print(which,'valid frac',ok.mean(),'x range',x0[ok].min(),x0[ok].max(),'y',y0[ok].min(),y0[ok].max())
    # spacing
    dx=np.hypot(np.diff(mx,axis=1),np.diff(my,axis=1)); dy=np.hypot(np.diff(mx,axis=0),np.diff(my,axis=0))
    print(' col spacing px: mean',np.nanmean(dx),' row spacing',np.nanmean(dy))
    addr=(y0*rig.width+x0)*3
    a0=(addr//8)*8  # first LDG.64 aligned
    line=a0//128
    R=(rows//8)*8; Cc=(cols//32)*32
    L=line[:R,:Cc]; O=ok[:R,:Cc]
    def count(shape):
        r,cw=shape
        t=L.reshape(R//r,r,Cc//cw,cw).transpose(0,2,1,3).reshape(-1,r*cw)
        o=O.reshape(R//r,r,Cc//cw,cw).transpose(0,2,1,3).reshape(-1,r*cw)
        t=np.where(o,t,-1)
        s=np.sort(t,axis=1)
        n=(np.diff(s,axis=1)!=0).sum(1)+1-(s[:,0]==-1)
        # sectors
        return n.mean()
    sec=a0//32
    for shape in [(1,32),(2,16),(4,8),(8,4)]:
        print(' warp shape rows x cols',shape,'lines/request',count(shape))
