
import sys, torch
sys.path.insert(0, "/root/repo")
from diffusion_model_project_b200 import engine, synth
from diffusion_model_project_b200.unet import B200UNet
torch.set_grad_enabled(False)
k = int(sys.argv[1]); f0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
m = B200UNet(**synth.UNET_KWARGS, device="cuda").load_state_dict(synth.synth_unet_state(seed=0))
st = m.build_program(4, 64, 64, fuse_small=False)
st["x_in"].hi.copy_(torch.randn(4, 1, 64, 64, 64, device="cuda").to(torch.bfloat16))
prog = st["program"]
names = [n for n, _ in prog.steps]
prog.run(torch.cuda.current_stream().cuda_stream); torch.cuda.synchronize()
chain = engine.Chain(prog, "cuda", max_ops=k, first_op=f0)
chain.run(torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print("OK", f0, k, names[f0:k])
