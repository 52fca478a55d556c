timeout 300 python -m pytest tests/test_gemm_tc_gpu.py -m gpu -q --timeout 60 2>&1 | tail -15
for pr in 0 1; do
timeout 120 python - <<PY
import torch, time
from kaldi_ctc_b200 import rnn
rnn.set_tuning("GEMM_PAIR", $pr)
for (tA,tB,M,N,K) in [(0,1,32000,1280,640),(0,0,32000,640,1280),(0,1,32000,1280,40)]:
    A=torch.randn((K,M) if tA else (M,K),device="cuda"); B=torch.randn((N,K) if tB else (K,N),device="cuda"); C=torch.zeros(M,N,device="cuda")
    flush=torch.empty(256<<20,dtype=torch.uint8,device="cuda")
    ts=[]
    for i in range(8):
        flush.zero_()
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); rnn.gemm(torch,tA,tB,M,N,K,1.0,A,A.shape[1],B,B.shape[1],0.0,C,N,math=rnn.MATH_TENSOR); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms=sorted(ts)[len(ts)//2]
    print("pair=%d %s M=%d N=%d K=%d: %.1f us  %.0f TFLOP/s (pair kernel used: %d)"%($pr,("T" if tA else "N")+("T" if tB else "N"),M,N,K,ms*1e3,2.0*M*N*K/ms/1e9,rnn.lib().b200rnnLastGemmUsedCtaPair()))
PY
done
