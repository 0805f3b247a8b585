import sys, os
sys.path[:0] = ["/root/repo/vm-placement-migration-gym_b200", "/root/repo/tests"]
import numpy as np, torch, ctypes as C
from vmgym import Config, VecVmEnv
from vmgym import _native as nv
from vmgym.ppo import PPOAgent, PPOConfig
torch.backends.cuda.matmul.allow_tf32 = False
kw = dict(pms=100, vms=300, arrival_rate=1.8182, service_length=100, training_steps=10000, eval_steps=100000, reward_function="wr")
N = 160
vec = VecVmEnv(Config(**kw), N, rng="philox")
vec.agent_step("firstfit", n_steps=150, want_action=False, want_valid=False)
torch.manual_seed(11)
a32 = PPOAgent(vec, PPOConfig(hidden_size=512, update_math="fp32"))
atc = PPOAgent(vec, PPOConfig(hidden_size=512, update_math="bf16", env_chunk=200))
atc._flat.copy_(a32._flat); atc.weights_changed()
obs = vec.observe().clone()
logits = a32.model.actor(obs).contiguous()
action, logprob, ent, mask = a32._heads(logits, -1.0, want_mask=True)
tc = atc._tc_network()
x = tc.cast_obs(obs)
a1, a2 = tc.hidden("actor", x)
h32 = a32.model.actor[:4](obs)
print("hidden max diff", (a2.float() - h32).abs().max().item())
lp = torch.empty((N, 300), device="cuda"); en = torch.empty((N, 300), device="cuda")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
A = a32.A
nv.check(nv.lib().vmgym_policy_fused(a2.data_ptr(), tc.head.w_pad.data_ptr(), tc.head.b_pad.data_ptr(), mask.data_ptr(), action.data_ptr(), N, 300, a32.A, 512, 0, 0, None, lp.data_ptr(), en.data_ptr(), st), "f")
torch.cuda.synchronize()
# reference per (env, vm)
A = a32.A
W = 4
m = ((mask.view(N, 300, W, 1) >> torch.arange(32, device="cuda", dtype=torch.int32)) & 1).bool().reshape(N, 300, 128)[:, :, :A]
z = logits.double().reshape(N, 300, A).masked_fill(m, -1e7)
lg = torch.log_softmax(z, -1)
rlp = lg.gather(-1, action.long().unsqueeze(-1)).squeeze(-1)
print("lp sum fused", lp.sum(1)[:4].tolist(), "ref", rlp.sum(1)[:4].tolist(), "heads", logprob[:4].tolist())
d = (lp.double() - rlp).abs()
print("max per-(env,vm) diff", d.max().item(), "at", np.unravel_index(int(d.argmax()), d.shape))
e, v = np.unravel_index(int(d.argmax()), d.shape)
print("row", e, v, "action", int(action[e, v]), "valid cols", (~m[e, v]).nonzero().flatten().tolist()[:10], "lp fused", lp[e, v].item(), "ref", rlp[e, v].item())
print("n bad", int((d > 1e-2).sum()), "of", d.numel())
P = 100
pl = vec.vm_placement.long()
n_empty = (pl == P + 1).sum(1); n_wait = (pl == P).sum(1); n_run = (pl < P).sum(1)
dd = (logprob.double() - rlp.sum(1))
print("diff/n_empty", (dd / n_empty.double())[:6].tolist(), "n_empty", n_empty[:6].tolist(), "n_wait", n_wait[:6].tolist())
allinv = m.all(-1)
print("rows all-invalid per env", allinv.sum(1)[:6].tolist(), "actions in all-invalid rows (env0)", action[0][allinv[0]].tolist()[:12])
print("ref lp of all-invalid rows env0", rlp[0][allinv[0]].tolist()[:6], "fused", lp[0][allinv[0]].tolist()[:6])
