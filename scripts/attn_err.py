import sys, torch
sys.path.insert(0, ".")
import dssmd_b200
torch.manual_seed(0)
B,C,H,W = 2,32,4,640
q,k,v,go = (torch.randn(B,C,H,W,device="cuda") for _ in range(4))
scale = C**-0.5
qq,kk,vv = (t.clone().requires_grad_(True) for t in (q,k,v))
out = dssmd_b200.epipolar_attention(qq,kk,vv,scale); out.backward(go)
q2,k2,v2 = (t.double().requires_grad_(True) for t in (q,k,v))
a,b,c = (t.permute(0,2,3,1).reshape(B*H,W,C) for t in (q2,k2,v2))
ref = torch.matmul(torch.softmax(torch.matmul(a,b.transpose(-1,-2))*scale,dim=-1),c).view(B,H,W,C).permute(0,3,1,2)
ref.backward(go.double())
def re(x,y): return float((x.double()-y).abs().max()/y.abs().max())
print("rel err vs float64: out %.2e dq %.2e dk %.2e dv %.2e" % (re(out,ref), re(qq.grad,q2.grad), re(kk.grad,k2.grad), re(vv.grad,v2.grad)))
