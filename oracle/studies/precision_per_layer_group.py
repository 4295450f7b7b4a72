"""TEST INFRASTRUCTURE ONLY -- CPU emulation of reduced-precision products (see profiles/r1_precision_study.txt)."""
import sys, numpy as np, torch, itertools
import os; ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT)
from oracle import deepfit_oracle as O
torch.set_grad_enabled(False)
name = sys.argv[1]
raw = np.load(os.path.join(ROOT, 'tests/golden/tutorial_clouds.npz'))[name]
p, bb = O.normalize_cloud(raw)
rng = np.random.default_rng(0)
q = rng.choice(len(p), 384, replace=False)
idx, dist = O.knn_ckdtree(p, q, 256); dist, idx = O.canonicalize_knn(dist, idx)
items = [O.dataset_item(p, int(c), 256, idx[j], dist[j]) for j,c in enumerate(q)]
P = torch.stack([it[0] for it in items]); Us = torch.stack([it[1] for it in items])
sd = O.seeded_state_dict(int(sys.argv[2]) if len(sys.argv)>2 else 0)
def fold(conv, bn):
    W = sd[conv+'.weight']; W = W.squeeze(-1) if W.dim()==3 else W
    b = sd[conv+'.bias']
    if bn is None: return W.clone(), b.clone()
    s = sd[bn+'.weight'] / torch.sqrt(sd[bn+'.running_var'] + 1e-5)
    return W * s[:,None], (b - sd[bn+'.running_mean']) * s + sd[bn+'.bias']
def bf16(x): return x.to(torch.bfloat16).to(torch.float32)
def split2(x):
    hi = bf16(x); lo = bf16(x - hi); return hi, lo
def mm(x, W, mode):
    if mode == 'fp32': return x @ W.t()
    if mode == 'b1': return bf16(x) @ bf16(W).t()
    xh, xl = split2(x); wh, wl = split2(W)
    if mode == 'b3': return xh @ wh.t() + xh @ wl.t() + xl @ wh.t()
    if mode == 'b2a': return xh @ wh.t() + xl @ wh.t()       # activations split, weights single
    if mode == 'b2w': return xh @ wh.t() + xh @ wl.t()       # weights split, activations single
    raise ValueError(mode)
def net(P, md):
    B,_,k = P.shape
    X = P.transpose(1,2)
    def trunk(pref, x, first_fp32, gm, fm):
        W,b = fold(pref+'.conv1', pref+'.bn1'); x = torch.relu((x @ W.t() if first_fp32 else mm(x, W, gm)) + b)
        W,b = fold(pref+'.conv2', pref+'.bn2'); x = torch.relu(mm(x, W, gm) + b)
        W,b = fold(pref+'.conv3', pref+'.bn3'); x = torch.relu(mm(x, W, gm) + b)
        g = x.max(dim=1)[0]
        W,b = fold(pref+'.fc1', pref+'.bn4'); g = torch.relu(mm(g, W, fm) + b)
        W,b = fold(pref+'.fc2', pref+'.bn5'); g = torch.relu(mm(g, W, fm) + b)
        W,b = fold(pref+'.fc3', None); return mm(g, W, fm) + b
    qv = trunk('feat.pointfeat.stn1', X, True, md['qstn'], md['fc']); qv = qv + torch.tensor([1.,0,0,0])
    R = O.quat_to_rotmat(qv); Xr = X @ R
    W,b = fold('feat.pointfeat.conv1', 'feat.pointfeat.bn1'); x = torch.relu(Xr @ W.t() + b)
    W,b = fold('feat.pointfeat.conv2', 'feat.pointfeat.bn2'); x = torch.relu(mm(x, W, md['pf']) + b)
    t2 = trunk('feat.pointfeat.stn2', x, False, md['stn'], md['fc'])
    t2 = (t2 + torch.eye(64).reshape(1,4096)).reshape(B,64,64)
    m = md['pf']
    if m == 'fp32': pf = torch.bmm(x, t2)
    elif m == 'b1': pf = torch.bmm(bf16(x), bf16(t2))
    else:
        xh, xl = split2(x); th, tl = split2(t2); pf = torch.bmm(xh, th) + torch.bmm(xh, tl) + torch.bmm(xl, th)
    W,b = fold('feat.conv2', 'feat.bn2'); y = torch.relu(mm(pf, W, md['enc']) + b)
    W,b = fold('feat.conv3', 'feat.bn3'); y = mm(y, W, md['enc']) + b
    g = y.max(dim=1)[0]
    W,b = fold('conv1', 'bn1')
    hg = mm(g, W[:, :1024], md['fc']) + b
    z = torch.relu(mm(pf, W[:, 1024:], md['h1']) + hg[:,None,:])
    W,b = fold('conv2', 'bn2'); z = torch.relu(mm(z, W, md['h2']) + b)
    W,b = fold('conv3', 'bn3'); z = torch.relu(mm(z, W, md['h3']) + b)
    W,b = fold('conv4', None); a = (z @ W.t() + b).squeeze(-1)
    return 0.01 + torch.sigmoid(a), R, Xr.transpose(1,2).contiguous()
w_ref, R_ref, t2_ref, pts_ref = O.weight_network(sd, P)
beta_ref, n_ref = O.fit_wjet(pts_ref, w_ref, 3); nw_ref = O.world_normals(n_ref, Us, R_ref); c_ref = O.principal_curvatures(beta_ref)
base = dict(qstn='b3', stn='b3', enc='b3', fc='b3', pf='b3', h1='b3', h2='b3', h3='b3')
def run(tag, **kw):
    md = dict(base); md.update(kw)
    w, R, pts = net(P, md)
    beta, n = O.fit_wjet(pts, w, 3); nw = O.world_normals(n, Us, R)
    ang = O.unoriented_angle_deg(nw.numpy(), nw_ref.numpy()); ce = O.rectified_rel_err(O.principal_curvatures(beta).numpy(), c_ref.numpy())
    print('%-28s w %.1e R %.1e | angle max %.2e p99 %.2e | curv max %.2e' % (tag, (w-w_ref).abs().max(), (R-R_ref).abs().max(), ang.max(), np.quantile(ang,0.99), ce.max()))
run('all b3')
for key in ['qstn','stn','enc','fc','pf','h1','h2','h3']:
    run(key+'=b1', **{key:'b1'})
for key in ['qstn','stn','enc','h2']:
    run(key+'=b2a', **{key:'b2a'}); run(key+'=b2w', **{key:'b2w'})
