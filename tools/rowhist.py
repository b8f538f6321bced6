import sys; sys.path.insert(0,'/root/repo')
import numpy as np
from dynoplan_b200 import testing
from dynoplan_b200.expander import GpuExpander
case=testing.synthetic_case("quad3d_v0",100000,4096,seed=0,delta=0.8,alpha=0.5,device=0)
ex=GpuExpander(case.robot,case.obstacles,case.library,device=0,capacity=1<<24)
got=ex.expand(case.nodes,case.radius,case.goal,case.goal_radius,0)
n=np.diff(got.row_ptr)
print("rows",len(n),"mean",n.mean(),"max",n.max())
print("hist", np.histogram(n,bins=[0,2,256,512,1024,2048,4096,8192,1<<20])[0].tolist())
print("pct", np.percentile(n,[1,10,50,90,99,99.9]).tolist())
