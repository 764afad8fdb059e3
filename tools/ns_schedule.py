# Optimal odd-quintic Newton-Schulz schedule (Polar-Express style) with safety margin above 1.
import numpy as np, scipy.optimize as so, sys
def optimal_quintic(l, u):
    xs=np.concatenate([np.geomspace(l,1,4000), np.linspace(1,u,200)])
    A=[];bb=[]
    for x in xs:
        A.append([ x, x**3, x**5,-1]); bb.append(1)
        A.append([-x,-x**3,-x**5,-1]); bb.append(-1)
    r=so.linprog([0,0,0,1],A_ub=np.array(A),b_ub=np.array(bb),bounds=[(None,None)]*3+[(0,None)],method="highs")
    a,b,c,e=r.x
    return (a,b,c),e
def schedule(l0,K,u):
    l=l0; out=[]
    for k in range(K):
        (a,b,c),e=optimal_quintic(l,u)
        s=1/(1+e)
        a,b,c=a*s,b*s,c*s
        # verify on fine grid
        xs=np.geomspace(l,u,200001); p=a*xs+b*xs**3+c*xs**5
        lo,hi=p.min(),p.max()
        out.append((a,b,c)); 
        print(f"// k={k}: domain [{l:.3e},{u}] -> [{lo:.6f},{hi:.6f}]", file=sys.stderr)
        l=lo
    return out
if __name__=="__main__":
    l0=float(sys.argv[1]); K=int(sys.argv[2]); u=float(sys.argv[3])
    for a,b,c in schedule(l0,K,u):
        print(f"    {{{a:.9f}f, {b:.9f}f, {c:.9f}f}},")
