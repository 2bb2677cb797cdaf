import sys, numpy as np
sys.path.insert(0, '/root/repo')
import latticefold_b200 as lf
from latticefold_b200 import synth
from oracle.pyoracle import Oracle
from tests.helpers import OracleOps
G=0
orc=Oracle(); ops=OracleOps(orc)
W,B,L,b,K,kappa,kind = 4, 1<<15, 5, 2, 15, 4, "scalar"
if len(sys.argv)>1: W,B,L,b,K,kappa=map(int,sys.argv[1:7]); kind=sys.argv[7]
prob=synth.make_instance(G,W,B,L,b,K,kappa,kind=kind,config_id=2,ops=ops)
ctx=lf.Context(G,0); pr=lf.NIFSProver(ctx,prob)
eproof,elc,ef,_=orc.nifs_prove(prob,orc.transcript(G))
proof,lc,f=pr.prove(prob,lf.Transcript(G))
ccs=prob["ccs"]; d=24; tau=3; s=ccs["s"]; t=ccs["t"]; l=ccs["l"]
secs=[("lin.msgs",s*(ccs["d"]+2)*d),("lin.v",tau*d),("lin.u",t*d)]
for h in ("dl","dr"):
    for k in range(K):
        secs += [(f"{h}.x[{k}]",(l+1)*d),(f"{h}.y[{k}]",kappa*d),(f"{h}.u[{k}]",t*d),(f"{h}.v[{k}]",tau*d)]
secs += [("fold.msgs",s*5*d),("fold.theta",2*K*tau*d),("fold.eta",2*K*t*d)]
o=0
for name,n in secs:
    if not np.array_equal(proof[o:o+n],eproof[o:o+n]):
        bad=np.nonzero(proof[o:o+n]!=eproof[o:o+n])[0]
        print("DIFF",name,"count",len(bad),"of",n,"first",bad[:6], proof[o+bad[0]], eproof[o+bad[0]])
    o+=n
assert o==proof.size
print("lc equal",np.array_equal(lc,elc),"f equal",np.array_equal(f,ef))
