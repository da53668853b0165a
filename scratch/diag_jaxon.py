import sys, numpy as np, torch
sys.path.insert(0,'tests'); sys.path.insert(0,'.'); sys.path.insert(0,'scikit-motionplan_b200')
from helpers import oracle_args, sample_q
from oracle import oracle
from skmp_b200.constraint import CollFreeConst
from skmp_b200.robot.jaxon import JaxonConfig
from skmp_b200.robot_model import Jaxon
from skmp_b200.sdf import Box, UnionSDF
from skmp_b200.rotmath import rotation_matrix
jaxon=Jaxon(); jaxon.reset_manip_pose()
env = UnionSDF([Box([4.0, 4.0, 0.2], pos=[0.0, 0.0, -0.1]), Box([0.6, 1.0, 0.8], pos=[0.9, 0.0, 0.4])])
held = Box([0.2, 0.3, 0.25]); held.translate([0.4, -0.3, 1.0]); held.rotate_with_matrix(rotation_matrix(0.4, [0.3, 1.0, -0.2]))
for name,kin in (("spheres",JaxonConfig().get_collision_kin()),("attached",JaxonConfig().get_attached_obstacle_kin(np.array([0.05,0,-0.12]),held))):
    rng=np.random.default_rng(31)
    qs=sample_q(kin,300,rng)*0.5; qs[:,33]+=1.0
    qs=qs.astype(np.float32).astype(np.float64)
    const=CollFreeConst(kin,env,jaxon)
    v,j=const.evaluate(torch.as_tensor(qs.astype(np.float32),device='cuda'),True)
    v=v.cpu().numpy(); j=j.cpu().numpy()
    tree,feat,jl,base=oracle_args(kin)
    rv,rj,kink=oracle.collfree(tree,qs,feat,jl,oracle.Sdf(env.primitives()),kin.get_radius_list(),base_type=base,with_kink=True)
    e=np.abs(j-rj)/np.maximum(1,np.abs(rj))
    e[kink<=2e-5]=0
    print(name,"max err",e.max(),"val err",np.abs(v-rv).max())
    print(" per-column max:",np.array2string(e.max(axis=(0,1)),precision=1))
    idx=np.unravel_index(e.argmax(),e.shape); print(" argmax",idx,"got",j[idx],"ref",rj[idx],"kink",kink[idx[0],idx[1]],"val",rv[idx[0],idx[1]])
    # map jacobian
    pts,jm=kin.map(torch.as_tensor(qs.astype(np.float32),device='cuda'))
    rp,rjm=oracle.solve_fk(tree,qs,feat,jl,base)
    em=np.abs(jm.cpu().numpy()-rjm)/np.maximum(1,np.abs(rjm))
    print(" map: pts err",np.abs(pts.cpu().numpy()-rp).max(),"jac err",em.max(),"per col",np.array2string(em.max(axis=(0,1,2)),precision=1))
    # error of sdf gradient: reconstruct via collfree with fp64
    v64,j64=const.evaluate(torch.as_tensor(qs,device='cuda'),True)
    e64=np.abs(j64.cpu().numpy()-rj)/np.maximum(1,np.abs(rj)); e64[kink<=1e-9]=0
    print(" fp64 jac err",e64.max())
    # how far are points from box: sdf gradient sensitivity
    big=np.argwhere(e>2e-5)
    print(" n entries >2e-5:",len(big), "distinct (cfg,row):",len({(a,b) for a,b,c in big}))
    for a,b,c in big[:5]:
        print("   ",a,b,c,j[a,b,c],rj[a,b,c],"kink",kink[a,b],"val",rv[a,b])
