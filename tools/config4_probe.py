import sys, os, json, numpy as np, torch
sys.path.insert(0, os.getcwd())
from evogp_b200.problem import Classification
from evogp_b200.tree import Forest, GenerateDescriptor
def ev(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts=[]
    for _ in range(reps):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True); a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))
rng=np.random.default_rng(0)
X=torch.from_numpy(rng.normal(size=(4096,13)).astype(np.float32)).cuda(); lab=torch.from_numpy(rng.integers(0,3,4096).astype(np.float32)).cuda()
d=GenerateDescriptor(max_tree_len=128,input_len=13,output_len=3,using_funcs=["+","-","*","/"],max_layer_cnt=7,const_samples=[-1,0,1],out_prob=0.5)
torch.manual_seed(0); f=Forest.random_generate(200000,d)
cls=Classification(datapoints=X,labels=lab,multi_output=True)
onehot=torch.nn.functional.one_hot(lab.long(),3).float().contiguous()
r={"fused_accuracy_ms":ev(lambda: cls.evaluate(f)),"sr_fitness_onehot_ms":ev(lambda: f.SR_fitness(X,onehot)),"unfused_ms":ev(lambda: cls.evaluate_unfused(f),reps=2,warm=1)}
a=cls.evaluate(f); b=cls.evaluate_unfused(f); r["trees_differing"]=int((a!=b).sum()); r["max_abs_diff"]=float((a-b).abs().max()); r["mean_acc"]=float(a.mean())
print(json.dumps(r))
