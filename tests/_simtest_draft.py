import sys, ctypes as C, numpy as np
sys.path.insert(0,'/root/repo/oracle'); sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo')
import pfft_oracle as po
import schedule_sim as ss
lib=C.CDLL('/root/repo/build/libplantest.so')

def run_case(kind, n, np_, flags=0, ni=None, no=None, howmany=1, sign=-1, kinds=None, skip=None, seed=0):
    d=len(n); ni_=ni or n; no_=no or n
    P=int(np.prod(np_))
    scheds=[ss.describe(lib, kind, n, np_, pid, flags, ni, no, howmany, None, None, sign, kinds, skip) for pid in range(P)]
    for s in scheds:
        assert s['error']=="", s['error']
    rng=np.random.default_rng(seed)
    r=len(np_)
    # global input
    if kind=='c2c':
        shape=list(ni_)+([howmany] if howmany>1 else [])
        xg=rng.standard_normal(shape)+1j*rng.standard_normal(shape)
    elif kind in('r2c','r2r'):
        shape=list(ni_)+([howmany] if howmany>1 else [])
        xg=rng.standard_normal(shape)
    else:
        # c2r: hermitian-consistent input from rfftn of random real data of size ni
        shape=list(ni_)+([howmany] if howmany>1 else [])
        xr=rng.standard_normal(shape)
        xg=np.fft.rfftn(xr, axes=list(range(d-1,-1,-1))[::-1]) if howmany==1 else None
        xg=np.fft.fftn(np.fft.rfft(xr,axis=d-1),axes=list(range(d-1)))
    tin=bool(flags&po.TRANSPOSED_IN); tout=bool(flags&po.TRANSPOSED_OUT)
    si=bool(flags&po.SHIFTED_IN); so=bool(flags&po.SHIFTED_OUT)
    padded=bool(flags&po.PADDED_R2C)
    user_in=[]
    for pid,s in enumerate(scheds):
        lni,lis=s['local_ni'],s['local_i_start']
        shift=[ni_[t]//2 if si else 0 for t in range(d)]
        ln=list(lni)
        if kind=='r2c':
            ln[-1]=ni_[-1]
        blk=po.extract_block(xg, ln, lis, None, shift)
        if kind=='r2c' and lni[-1]!=ni_[-1]:
            pad=[(0,0)]*blk.ndim; pad[d-1]=(0,lni[-1]-ni_[-1]); blk=np.pad(blk,pad,constant_values=np.nan)
        if tin:
            order=po.mem_order(d,r,True)
            blk=np.ascontiguousarray(np.transpose(blk, order+list(range(d,blk.ndim))))
        user_in.append(blk.reshape(-1))
    outs=ss.simulate(scheds,user_in)
    # oracle
    xin = xg
    want=po.global_transform({'c2c':po.C2C,'r2c':po.R2C,'c2r':po.C2R,'r2r':po.R2R}[kind], xin, n, ni, no, sign, flags, kinds, skip if skip is None else [skip[min(t,r)] for t in range(d)])
    err=0
    for pid,s in enumerate(scheds):
        lno,los=s['local_no'],s['local_o_start']
        shift=[no_[t]//2 if so else 0 for t in range(d)]
        ln=list(lno)
        order=po.mem_order(d,r,True) if tout else list(range(d))
        shp=[lno[t] for t in order]+([howmany] if howmany>1 else [])
        got=outs[pid].reshape(shp)
        if kind=='c2r' and lno[-1]!=no_[-1]:
            got=got[...,:no_[-1]] if howmany==1 else got[...,:no_[-1],:]
            ln[-1]=no_[-1]
        inv=np.argsort(order).tolist()
        got=np.transpose(got, inv+list(range(d,got.ndim)))
        ref=po.extract_block(want, ln, los, None, shift)
        if ref.size:
            err=max(err, np.abs(got-ref).max()/max(1e-300,np.abs(want).max()))
    return err, len(scheds[0]['stages'])

T_IN,T_OUT=po.TRANSPOSED_IN,po.TRANSPOSED_OUT
cases=[
 ('c2c',[29,27,31],[2,2],0),('c2c',[29,27,31],[2,2],T_OUT),('c2c',[29,27,31],[2,2],T_IN),
 ('c2c',[8,8,8],[1,1],0),('c2c',[8,6,4],[1,1],T_OUT),('c2c',[8,6,4],[1,1],T_IN),
 ('c2c',[16,12,10],[4],0),('c2c',[16,12,10],[4],T_OUT),('c2c',[16,12,10],[3],T_IN),
 ('c2c',[13,14,19,17],[2,2,2],T_OUT),('c2c',[13,14,19,17],[2,2,2],0),('c2c',[13,14,19,17],[2,2,2],T_IN),
 ('c2c',[13,14,19,17],[2,2],0),('c2c',[13,14,19,17],[3,2],T_OUT),
 ('c2c',[5,4,3],[3,2],0),('c2c',[4,4,4],[3,3],T_OUT),
 ('r2c',[29,27,31],[2,2],0),('r2c',[29,27,31],[2,2],T_OUT),('r2c',[16,12,10],[2,2],T_OUT|po.PADDED_R2C),
 ('c2r',[29,27,31],[2,2],0),('c2r',[29,27,30],[2,2],T_IN),('c2r',[16,12,10],[2,2],T_IN|po.PADDED_R2C),
 ('r2c',[8,6,10],[1,1],0),('c2r',[8,6,10],[1,1],0),
]
for c in cases:
    kind,n,np_,fl=c
    for sign in ((-1,1) if kind=='c2c' else ((-1,) if kind=='r2c' else (1,))):
        e,ns=run_case(kind,n,np_,fl,sign=sign)
        print(kind,n,np_,fl,sign,'stages',ns,'err %.2e'%e, 'OK' if e<1e-12 else 'FAIL')
