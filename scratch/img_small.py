import torch, sys
sys.path.insert(0, '.')
from dns_slam_b200 import _lib
dev = torch.device("cuda:0")
for rows in (94000, 250047, 23500):
    for M, N, RS in [(80, 64, 128), (112, 64, 128), (33, 32, 128), (3, 32, 128)]:
        A = torch.randn(rows, M, device=dev); B = torch.randn(rows, N, device=dev); C = torch.zeros(M, N, device=dev)
        _lib.check(_lib.lib().dns_debug_gemm_img(_lib.ptr(A), M, M, _lib.ptr(B), N, N, rows, RS, _lib.ptr(C), _lib.stream()))
        torch.cuda.synchronize()
