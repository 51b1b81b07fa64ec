import sys, torch, numpy as np
sys.path.insert(0,'.')
import bench
from hsr_env_b200.env import BatchedHSREnv
from hsr_env_b200.spaces import Box
from hsr_env_b200.util import GoalSpec
dev=torch.device('cuda',0); n=4096
goals=[GoalSpec(a=Box(bench.BLOCK_LO,bench.BLOCK_HI),b=Box(bench.GOAL_LO,bench.GOAL_HI),distance=bench.GEOFENCE)]
env=BatchedHSREnv(bench.BLOB,goals,steps_per_action=300,n_envs=n,device=dev,seed=0)
lo=torch.tensor(env.model.act_ctrlrange[:,0],dtype=torch.float32,device=dev); hi=torch.tensor(env.model.act_ctrlrange[:,1],dtype=torch.float32,device=dev)
gen=torch.Generator(device=dev).manual_seed(0)
flush=torch.empty(256*1024*1024,dtype=torch.uint8,device=dev)
env.reset(); done=torch.zeros(n,dtype=torch.bool,device=dev)
ts=[]
for k in range(60):
    a=lo+(hi-lo)*torch.rand(n,env.nu,generator=gen,device=dev)
    if k%2==0: flush.fill_(k&0xff)
    env.reset(mask=done)
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record(); obs,r,done,inf=env.step(a); e1.record(); torch.cuda.synchronize()
    ts.append((e0.elapsed_time(e1), 'F' if k%2==0 else '-'))
print(" ".join("%.1f%s"%t for t in ts))
