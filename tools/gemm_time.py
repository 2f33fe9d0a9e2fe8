import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepfake_detection_b200 import _lib
def t(M,N,K,stats=True, reps=20):
    A=torch.randn(M,K,device='cuda').bfloat16(); B=torch.randn(N,K,device='cuda').bfloat16(); C=torch.empty(M,N,device='cuda',dtype=torch.bfloat16)
    s1=torch.zeros(8,N,dtype=torch.float64,device='cuda'); s2=torch.zeros_like(s1)
    st=torch.cuda.current_stream().cuda_stream
    for _ in range(3): _lib.call("dfd_gemm_tn",A.data_ptr(),B.data_ptr(),C.data_ptr(),M,N,K,0,s1.data_ptr() if stats else None,s2.data_ptr() if stats else None,None,st)
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): _lib.call("dfd_gemm_tn",A.data_ptr(),B.data_ptr(),C.data_ptr(),M,N,K,0,s1.data_ptr() if stats else None,s2.data_ptr() if stats else None,None,st)
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/reps
    print("DBG=%s M=%d N=%d K=%d stats=%s ms=%.3f GB/s=%.0f"%(os.environ.get("DFD_DBG","0"),M,N,K,stats,ms,2*(M*K+M*N)/ms/1e6))
shapes=[(3211264,96,16),(802816,144,24),(200704,240,40),(3211264,16,96),(50176,672,112),
        (802816,256,64),(802816,64,256),(200704,512,128),(50176,1024,256),(12544,2048,512)]      # + resnet50 1x1 layers
if os.environ.get("GT_ONE"): shapes=shapes[:1]+shapes[3:4]
for shp in shapes:
    t(*shp); t(*shp, stats=False)
