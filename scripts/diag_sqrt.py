import torch
x = torch.rand(1<<20) * 10
a = torch.sqrt(x); b = torch.sqrt(x.cuda()).cpu()
print("torch sqrt cpu-vs-cuda mismatches:", int((a!=b).sum()), "of", x.numel())
