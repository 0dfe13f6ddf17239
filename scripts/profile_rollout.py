import sys, torch
sys.path.insert(0, "/root/repo")
from torch.profiler import ProfilerActivity, profile
from molecular_dynamics_neural_operator_b200 import synthetic as syn
from molecular_dynamics_neural_operator_b200.graph_kernel import KernelNN, construct_pairdata
wl = syn.WORKLOADS["synthetic_10k"]; dev = torch.device("cuda:0"); torch.manual_seed(0)
model = KernelNN(wl.width, wl.ker_width, wl.depth, 6, in_width=7, out_width=3, window_size=2, num_residues=wl.num_atoms, precision="bf16").to(dev).eval()
traj = torch.from_numpy(syn.trajectory(wl, 2, seed=0)).to(dev); aa = torch.from_numpy(syn.residue_ids(wl)).to(dev)
data = construct_pairdata(traj, aa, threshold=wl.cutoff, device=dev)
def step(data):
    out = model(data, single_example=True)
    sp = data.x_position[-1] + 0.05 * torch.tanh(torch.nan_to_num(out))
    return construct_pairdata(torch.cat([data.x_position[1:], sp[None]]), aa, threshold=wl.cutoff, device=dev)
with torch.no_grad():
    for _ in range(3): data = step(data)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3): data = step(data)
        torch.cuda.synchronize()
rows=[(e.device_time_total/3e3, e.count//3, e.key[:100]) for e in prof.key_averages() if e.device_time_total and str(e.device_type).endswith("CUDA")]
rows.sort(reverse=True); print("device ms/step", sum(r[0] for r in rows))
for r in rows[:14]: print("%9.3f ms n=%4d %s" % r)
print(prof.key_averages().table(sort_by="cpu_time_total", row_limit=12, max_name_column_width=60))
