"""Accuracy of the hybrid tensor-core spectrum in its REAL operand layout (profiles/r02_tensor_core_hybrid_study.md):
integer pre-emphasis 100 x[n] - 97 x[n-1] split exactly into two fp16 pieces, stage-1 matrix (Povey window, 32-point DFT
over n1, inter-stage twiddle) per residue class in two fp16 pieces, frame sum in three spare K slots, K = 16 accumulation
steps in the order the kernel would issue them (small products first), fp32 accumulation with round-to-nearest or
truncation; stage 2 as an fp32 FFT.  Prints max |log-mel error| vs the fp64 oracle / vs the fp32 oracle per signal class.

CPU only (test infrastructure, like oracle/): never imported by the product.
"""
import sys, numpy as np
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import fbank as F, signals
mel64 = F.mel_banks(80, 512, 16000.0, 20.0, 0.0, np.float32).astype(np.float64)
w = F.povey_window(400, np.float64)
f16=lambda a: a.astype(np.float16).astype(np.float64)
def trunc32(v):
    v32=v.astype(np.float32); bad=np.abs(v32.astype(np.float64))>np.abs(v); v32[bad]=np.nextafter(v32[bad],np.float32(0)); return v32.astype(np.float64)
rn32=lambda v: v.astype(np.float32).astype(np.float64)
# G without pre-emphasis: per class n2: rows n1 (25), cols 32
def Gclass(n2):
    g=np.zeros((25,32)); 
    for n1 in range(25):
        n=16*n1+n2
        if n>=400: continue
        for k1 in range(17):
            t=w[n]*np.exp(-2j*np.pi*n1*k1/32.0)
            if k1!=16: t=t*np.exp(-2j*np.pi*n2*k1/512.0)
            if k1==0: g[n1,0]=t.real
            elif k1==16: g[n1,1]=t.real
            else: g[n1,2*k1]=t.real; g[n1,2*k1+1]=t.imag
    return g
def run(wave, npieces_g, prods, rnd, order_small_first=True, f32in=False):
    x=F.frames_of(wave.astype(np.float64)); m=x.shape[0]
    xp=np.concatenate([x[:,:1],x[:,:-1]],axis=1)
    if not f32in:
        v=100*x-97*xp   # exact int, |v|<2^24 ; scale 1/100 folded into G
        gs=1/100.0
    else:
        v=rn32(x-rn32(0.97*xp)); gs=1.0
    hi=f16(v/256.0)*256.0; lo=f16((v-hi)) ; res=v-hi-lo
    S=x.sum(axis=1); 
    s1=f16(S/4096.0)*4096.0; s2=f16(S-s1); s3=S-s1-s2
    Y=np.zeros((m,16,32))
    maxres=np.abs(res).max()
    for n2 in range(16):
        g=Gclass(n2)*gs
        # DC column: z -= w[n]*0.03*mu  -> coefficient on S: -0.03/400*sum_n1 w*T
        vdc=-(0.03/400.0)*(Gclass(n2).sum(axis=0))
        # wait: z[n]=w[n]*(y[n]-0.03 mu); n=0 has w=0
        pieces=[]; r=g.copy()
        for i in range(npieces_g):
            p=f16(r*2.0**(11*i))/2.0**(11*i); pieces.append(p); r=r-p
        vp=[]; r=vdc.copy()
        for i in range(npieces_g):
            p=f16(r*2.0**(11*i)*4096)/(2.0**(11*i)*4096); vp.append(p); r=r-p
        idx=16*np.arange(25)+n2; ok=idx<400; idx=np.minimum(idx,399)
        ah=hi[:,idx]*ok; al=lo[:,idx]*ok
        # k-steps: plane0 a=0,1 (n1 0..4,10..14) ; plane0 a=2 (n1 20..24); plane1 a=0,1 (n1 5..9, 15..19)
        steps=[[0,1,2,3,4,10,11,12,13,14],[20,21,22,23,24],[5,6,7,8,9,15,16,17,18,19]]
        terms=[]
        for (pa,pg) in prods:
            A=ah if pa==0 else al
            for si,st in enumerate(steps):
                contrib=A[:,st]@pieces[pg][st]
                if si==0 and pa==0:
                    # DC slots ride in plane0 a=0 of the hi plane
                    sv = s1[:,None]*vp[pg][None] + s2[:,None]*vp[pg][None] + (s3[:,None]*vp[pg][None] if pg==0 else 0)
                    contrib=contrib+sv
                terms.append((pa+pg,contrib))
        if order_small_first: terms.sort(key=lambda t:-t[0])
        acc=np.zeros((m,32))
        for _,c in terms: acc=rnd(acc+c)
        Y[:,n2,:]=acc
    Yc=np.zeros((m,16,17),np.complex128)
    Yc[:,:,0]=Y[:,:,0]; Yc[:,:,16]=Y[:,:,1]*np.exp(-2j*np.pi*np.arange(16)*16/512.0)[None]
    for k1 in range(1,16): Yc[:,:,k1]=Y[:,:,2*k1]+1j*Y[:,:,2*k1+1]
    X=np.fft.fft(Yc.astype(np.complex64),axis=1).astype(np.complex64).transpose(0,2,1)
    power=np.zeros((m,257))
    for k1 in range(17):
        for k2 in range(16):
            k=k1+32*k2; kk=k if k<=256 else 512-k
            if k1 in (0,16) and k>256: continue
            power[:,kk]=np.abs(X[:,k1,k2].astype(np.complex128))**2
    melp=np.concatenate([mel64,np.zeros((80,1))],axis=1)
    return np.log(np.maximum(power@melp.T,float(F.EPS_F32))), maxres
P3=[(0,0),(0,1),(1,0)]; P4=[(0,0),(0,1),(1,0),(1,1)]; P5=[(0,0),(0,1),(0,2),(1,0),(1,1)]
for kind in ('dcsine','white','speech','lsb','square'):
    wv=signals.make(kind,8000,7); ref=F.fbank(wv.astype(np.float64),dtype=np.float64,mel=mel64)
    ref32=F.fbank(wv.astype(np.float32),mel=mel64.astype(np.float32)).astype(np.float64)
    out=[]
    for name,(ng,pr,rnd,osf) in {'P3 rn':(2,P3,rn32,True),'P3 trunc':(2,P3,trunc32,True),'P3 trunc bigfirst':(2,P3,trunc32,False),'P4 rn':(2,P4,rn32,True),'P5 rn':(3,P5,rn32,True),'P5 trunc':(3,P5,trunc32,True)}.items():
        got,mr=run(wv,ng,pr,rnd,osf)
        out.append('%s %.2e/%.2e'%(name,np.abs(got-ref).max(),np.abs(got-ref32).max()))
    print(kind,'fp32ref %.2e'%np.abs(ref32-ref).max(),' | '.join(out), 'maxres',mr)
