import os, sys, time, ctypes, torch
sys.path.insert(0, '/root/repo')
from evogp_b200 import _native
from evogp_b200.tree import Forest, GenerateDescriptor
torch.manual_seed(0)
P,L,N,V=100000,64,1024,3
d=GenerateDescriptor(max_tree_len=L,input_len=V,output_len=1,using_funcs=["+","-","*","/"],max_layer_cnt=6,const_samples=[-1,0,1])
f=Forest.random_generate(P,d)
X=torch.rand(N,V,device='cuda')*2-1; y=(X[:,:1]**2).contiguous()
hv,ht,hs=(a.cpu().pin_memory() for a in (f.batch_node_value,f.batch_node_type,f.batch_subtree_size))
hX,hy=X.cpu().pin_memory(),y.cpu().pin_memory(); out=torch.empty(P).pin_memory()
vp=lambda t: ctypes.c_void_p(t.data_ptr())
abi=_native.abi()
def call(): _native.check(abi.evogp_SR_fitness_host(P,N,L,V,1,1,vp(hv),vp(ht),vp(hs),vp(hX),vp(hy),vp(out),0),"host")
for _ in range(5): call()
t0=time.perf_counter()
for _ in range(30): call()
dt=(time.perf_counter()-t0)/30
print(os.environ.get("EVOGP_HOST_CHUNKS","8"), f"{dt*1e3:.3f} ms/step  {P*N/dt:.3e} tree-evals/s  eff H2D {38.6e6/dt/1e9:.1f} GB/s")
# plain pinned copy bandwidth for reference
big=torch.empty(64*1024*1024,dtype=torch.uint8).pin_memory(); dev=torch.empty_like(big,device='cuda')
torch.cuda.synchronize(); t0=time.perf_counter()
for _ in range(10): dev.copy_(big,non_blocking=True)
torch.cuda.synchronize(); print("pinned H2D 64 MiB:", 64*1024*1024*10/(time.perf_counter()-t0)/1e9, "GB/s")
