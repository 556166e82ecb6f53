import importlib, sys, torch
sys.path.insert(0, "/root/repo")
pkg = importlib.import_module("xai-audio-deepfakes_b200"); H = pkg.hifigan
from oracle import vocoder as V
bf16r = lambda t: t.to(torch.bfloat16).float()
def rel(a,b): a,b=a.float().cpu(),b.float().cpu(); return float((a-b).norm()/b.norm())
for std in (0.01, 0.02, 0.03):
    W = H.init_weights(H.HifiganConfig, seed=1, std=std)
    gen = H.HifiganGenerator(W)
    g = torch.Generator().manual_seed(2)
    mel = -4 + 2*torch.randn(2,80,9,generator=g)
    wav = gen.decode_batch(mel)
    Wq = {k:(bf16r(v) if k.endswith("weight") and not k.startswith("conv_post") else v) for k,v in W.items()}
    ref = V.generator(mel, Wq); refq = V.generator(mel, Wq, quantize=bf16r); ref0 = V.generator(mel, W)
    print(std, "max|ref|", float(ref.abs().max()), "vs fp32(bf16 w)", rel(wav,ref), "vs quantized oracle", rel(wav,refq), "oracle q vs exact", rel(refq,ref), "bf16-weights vs fp32-weights", rel(ref, ref0))
