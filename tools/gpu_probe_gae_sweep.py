import torch, sys
sys.path.insert(0, ".")
from cadre_b200 import ppo
dev="cuda:0"
for (E,T) in ((65536,1024),(65536,800),(65536,512)):
    r=torch.rand(E,T+1,device=dev); v=torch.randn(E,T+1,device=dev); m=(torch.rand(E,T+1,device=dev)>0.02).float(); nv=torch.randn(E,device=dev)
    ret=torch.empty(E,T+1,device=dev); adv=torch.empty(E,T,device=dev)
    for _ in range(3): ppo.gae(r,v,m,nv,ret,adv)
    torch.cuda.synchronize(); a,b=torch.cuda.Event(True),torch.cuda.Event(True); a.record()
    for _ in range(10): ppo.gae(r,v,m,nv,ret,adv)
    b.record(); torch.cuda.synchronize(); ms=a.elapsed_time(b)/10
    print(E,T,round(ms,4),"ms",round(20.0*E*T/ms/1e6,1),"GB/s", round(20.0*E*T/ms/1e6/6453.1,3))
    del r,v,m,ret,adv
