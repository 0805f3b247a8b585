"""torch.profiler breakdown of DRLVMPAgent.act at S1000 (1024 envs)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "vm-placement-migration-gym_b200"))
import torch
from torch.profiler import profile, ProfilerActivity
from vmgym import Config, VecVmEnv
from vmgym.drlvmp import DRLVMPAgent, DRLVMPConfig
kw = dict(pms=1000, vms=3000, arrival_rate=1.6, service_length=1000, training_steps=10000, eval_steps=100000, seed=0,
          reward_function="wr", sequence="highuniform", allow_null_action=True)
vec = VecVmEnv(Config(**kw), 1024, rng="philox")
vec.agent_step("bestfit", 3000, want_obs=False, want_action=False, want_valid=False)
torch.set_float32_matmul_precision("high")
agent = DRLVMPAgent(vec, DRLVMPConfig(hidden_size=512)); agent.eval()
obs = vec.observe()
agent.act(obs)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    agent.act(obs)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
