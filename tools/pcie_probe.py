"""tools/pcie_probe.py -- what the host link of this box sustains (pinned <-> device copies), to place bench.py e2e."""
import torch, time
n=4096*4096*3
h=[torch.empty(n,dtype=torch.uint8).pin_memory() for _ in range(4)]
d=[torch.empty(n,dtype=torch.uint8,device='cuda') for _ in range(4)]
ho=[torch.empty(n//3,dtype=torch.uint8).pin_memory() for _ in range(4)]
do=[torch.empty(n//3,dtype=torch.uint8,device='cuda') for _ in range(4)]
s1,s2=torch.cuda.Stream(),torch.cuda.Stream()
def run(both,reps=32):
    torch.cuda.synchronize(); t=time.time()
    for i in range(reps):
        with torch.cuda.stream(s1): d[i%4].copy_(h[i%4],non_blocking=True)
        if both:
            with torch.cuda.stream(s2): ho[i%4].copy_(do[i%4],non_blocking=True)
    torch.cuda.synchronize(); dt=time.time()-t
    return reps*n/dt/1e9
for both in (False,True,False,True): print("d2h concurrent" if both else "h2d only", round(run(both),2),"GB/s H2D")
