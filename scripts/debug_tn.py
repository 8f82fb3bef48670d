import sys, torch
sys.path.insert(0, '.')
from sgs_gnn_b200 import ops
from sgs_gnn_b200._lib import lib
dev = torch.device('cuda:0')
for (k, m, n) in [(32, 128, 128), (4096, 256, 128)]:
    a1 = torch.zeros(k, m); a1[0, 5] = 1.0; a1[1, 37] = 3.0
    b1 = torch.zeros(k, n); b1[0, 3] = 2.0; b1[1, 70] = 5.0
    out = torch.full((m, n), -7.0, device=dev)
    got = ops.gemm(a1.to(dev), 1, m, b1.to(dev), 1, n, m, n, k, out=out, precision=ops.PREC_TF32)
    torch.cuda.synchronize()
    print(k, m, n, 'err:', lib().sgs_last_error())
    got = got.cpu()
    print(' count(-7):', int((got == -7.0).sum()), 'of', m * n, ' nonzero(not -7, not 0):', torch.nonzero((got != -7.0) & (got != 0)).tolist()[:10],
          got[(got != -7.0) & (got != 0)][:10].tolist())
    print(' expect (5,3)=2, (37,70)=15')
