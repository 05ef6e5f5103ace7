"""How much of the forward's face loop does useful work?  CPU only (tape from the kernel emulator, tests/emu): for a
sample of warps of a scene with the bench's rays/points ratio, per warp iteration and 4-face chunk: which lanes still
have faces (utilisation), which chunks contain a front face (dp > 0) for some lane (skippable otherwise).

    python tests/tools/face_loop_stats.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
import numpy as np  # noqa: E402

import common  # noqa: E402
import emu  # noqa: E402

n,width,height=50000,416,240
case = common.scene_case(num_points=n, width=width, height=height, q=2); f=case.foam
pipe = emu.EmuPipeline(3)
for _ in range(2):
    out = pipe.trace_forward(f.points, f.attributes, f.adjacency, f.offsets, case.rays, case.start, case.quantiles, scene_version=1, record_tape=True)
cells, t1, count = emu.tape_records(pipe, height, width)
W,S,_=cells.shape
# ray dirs per (warp, lane)
dirs=np.zeros((W,32,3),np.float32); valid_lane=np.zeros((W,32),bool)
bx_n=(width+15)//16
for w in range(W):
    block,warp=divmod(w,4); bx,by=block%bx_n, block//bx_n
    for lane in range(32):
        x=bx*16+(warp&1)*8+(lane&7); y=by*8+(warp>>1)*4+(lane>>3)
        if x<width and y<height:
            d=case.rays[y,x,3:]; dirs[w,lane]=d/np.linalg.norm(d); valid_lane[w,lane]=True
off=f.offsets.astype(np.int64); adj=f.adjacency
diff=(f.points[adj]-np.repeat(f.points,np.diff(off),axis=0)).astype(np.float16).astype(np.float32)
nf_all=np.diff(off)
maxf=int(nf_all.max()); nch=(maxf+3)//4
rng=np.random.default_rng(0)
ws=rng.choice(W,size=300,replace=False)
tot_chunks=0; needed_chunks=0; lane_chunks=0; lane_needed=0; faces_tot=0; faces_front=0
for w in ws:
    for k in range(S):
        c=cells[w,k]; m=(c!=0xFFFFFFFF)
        if not m.any(): break
        lanes=np.nonzero(m)[0]
        need=np.zeros((len(lanes),nch),bool); have=np.zeros((len(lanes),nch),bool)
        for i,l in enumerate(lanes):
            b=off[c[l]]; nf=nf_all[c[l]]
            dp=diff[b:b+nf]@dirs[w,l]
            front=dp>0
            faces_tot+=nf; faces_front+=int(front.sum())
            pad=np.zeros(nch*4,bool); pad[:nf]=front
            need[i]=pad.reshape(nch,4).any(axis=1)
            have[i,: (nf+3)//4]=True
        it=have.any(axis=0).sum()           # chunk iterations the warp executes
        tot_chunks+=it
        needed_chunks+=(need.any(axis=0)).sum()
        lane_chunks+=have.sum(); lane_needed+=need.sum()
print('front-face fraction', faces_front/faces_tot)
print('warp chunk iterations', tot_chunks, 'needed by some lane', needed_chunks, 'skippable', 1-needed_chunks/tot_chunks)
print('per-lane chunks', lane_chunks, 'needed', lane_needed, 'lane-level skippable', 1-lane_needed/lane_chunks, 'lane utilisation of executed iterations', lane_chunks/(tot_chunks*32))
print('useful (lane, chunk) slots of all executed slots', lane_needed/(tot_chunks*32))
