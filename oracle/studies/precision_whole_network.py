"""TEST INFRASTRUCTURE ONLY -- CPU emulation of reduced-precision products (see profiles/r1_precision_study.txt)."""
import sys, numpy as np, torch
import os; ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT)
from oracle import deepfit_oracle as O
torch.set_grad_enabled(False)
raw = np.load(os.path.join(ROOT, 'tests/golden/tutorial_clouds.npz'))[sys.argv[1]]
p, bb = O.normalize_cloud(raw)
rng = np.random.default_rng(0)
q = rng.choice(len(p), 512, replace=False)
idx, dist = O.knn_ckdtree(p, q, 256); dist, idx = O.canonicalize_knn(dist, idx)
items = [O.dataset_item(p, int(c), 256, idx[j], dist[j]) for j,c in enumerate(q)]
P = torch.stack([it[0] for it in items]); Us = torch.stack([it[1] for it in items])
sd = O.seeded_state_dict(int(sys.argv[2]) if len(sys.argv)>2 else 0)

def fold(sd, conv, bn):
    W = sd[conv+'.weight'].squeeze(-1) if sd[conv+'.weight'].dim()==3 else sd[conv+'.weight']
    b = sd[conv+'.bias']
    if bn is None: return W.clone(), b.clone()
    s = sd[bn+'.weight'] / torch.sqrt(sd[bn+'.running_var'] + 1e-5)
    return W * s[:,None], (b - sd[bn+'.running_mean']) * s + sd[bn+'.bias']

def bf16(x): return x.to(torch.bfloat16).to(torch.float32)
def tf32(x):
    i = x.view(torch.int32); i = (i + 0x1000) & ~0x1FFF  # round to 10 bits mantissa (approx RN)
    return i.view(torch.float32)
def split2(x):  # hi + lo bf16
    hi = bf16(x); lo = bf16(x - hi); return hi, lo

def lin(x, W, b, mode):
    # x [..., K], W [N,K]
    if mode == 'fp32': return x @ W.t() + b
    if mode == 'bf16': return bf16(x) @ bf16(W).t() + b
    if mode == 'tf32': return tf32(x) @ tf32(W).t() + b
    if mode == 'bf16x3':
        xh, xl = split2(x); wh, wl = split2(W)
        return xh @ wh.t() + xh @ wl.t() + xl @ wh.t() + b
    if mode == 'bf16x2':  # activations split only (weights hi only)?? no: hi*hi + lo*hi + hi*lo without lo*lo = x3. x2 = act exact-ish, weight bf16
        xh, xl = split2(x); wh = bf16(W)
        return xh @ wh.t() + xl @ wh.t() + b
    raise ValueError(mode)

def net(P, mode_pt, mode_fc):
    B,_,k = P.shape
    X = P.transpose(1,2)  # [B,k,3]
    def trunk(pref, x, first_mode):
        W,b = fold(sd, pref+'.conv1', pref+'.bn1'); x = torch.relu(lin(x, W, b, first_mode))
        W,b = fold(sd, pref+'.conv2', pref+'.bn2'); x = torch.relu(lin(x, W, b, mode_pt))
        W,b = fold(sd, pref+'.conv3', pref+'.bn3'); x = torch.relu(lin(x, W, b, mode_pt))
        g = x.max(dim=1)[0]
        W,b = fold(sd, pref+'.fc1', pref+'.bn4'); g = torch.relu(lin(g, W, b, mode_fc))
        W,b = fold(sd, pref+'.fc2', pref+'.bn5'); g = torch.relu(lin(g, W, b, mode_fc))
        W,b = fold(sd, pref+'.fc3', None); return lin(g, W, b, mode_fc)
    qv = trunk('feat.pointfeat.stn1', X, 'fp32'); qv = qv + torch.tensor([1.,0,0,0])
    R = O.quat_to_rotmat(qv)
    Xr = X @ R
    W,b = fold(sd, 'feat.pointfeat.conv1', 'feat.pointfeat.bn1'); x = torch.relu(lin(Xr, W, b, 'fp32'))
    W,b = fold(sd, 'feat.pointfeat.conv2', 'feat.pointfeat.bn2'); x = torch.relu(lin(x, W, b, mode_pt))
    t2 = trunk('feat.pointfeat.stn2', x, mode_pt)
    t2 = (t2 + torch.eye(64).reshape(1,4096)).reshape(B,64,64)
    if mode_pt == 'fp32': pf = torch.bmm(x, t2)
    elif mode_pt == 'bf16': pf = torch.bmm(bf16(x), bf16(t2))
    elif mode_pt == 'tf32': pf = torch.bmm(tf32(x), tf32(t2))
    else:
        xh, xl = split2(x); th, tl = split2(t2); pf = torch.bmm(xh, th) + torch.bmm(xh, tl) + torch.bmm(xl, th)
    W,b = fold(sd, 'feat.conv2', 'feat.bn2'); y = torch.relu(lin(pf, W, b, mode_pt))
    W,b = fold(sd, 'feat.conv3', 'feat.bn3'); y = lin(y, W, b, mode_pt)
    g = y.max(dim=1)[0]
    W,b = fold(sd, 'conv1', 'bn1')
    hg = lin(g, W[:, :1024], b, mode_fc)   # [B,512]
    z = torch.relu(lin(pf, W[:, 1024:], torch.zeros(512), mode_pt) + hg[:,None,:])
    W,b = fold(sd, 'conv2', 'bn2'); z = torch.relu(lin(z, W, b, mode_pt))
    W,b = fold(sd, 'conv3', 'bn3'); z = torch.relu(lin(z, W, b, mode_pt))
    W,b = fold(sd, 'conv4', None); a = lin(z, W, b, 'fp32').squeeze(-1)
    w = 0.01 + torch.sigmoid(a)
    return w, R, t2, Xr.transpose(1,2).contiguous(), a

w_ref, R_ref, t2_ref, pts_ref = O.weight_network(sd, P)
beta_ref, n_ref = O.fit_wjet(pts_ref, w_ref, 3)
nw_ref = O.world_normals(n_ref, Us, R_ref)
c_ref = O.principal_curvatures(beta_ref)
for mode_pt, mode_fc in [('fp32','fp32'), ('bf16','bf16'), ('bf16','fp32'), ('tf32','tf32'), ('bf16x3','bf16x3'), ('bf16x2','bf16x2')]:
    w, R, t2, pts, a = net(P, mode_pt, mode_fc)
    beta, n = O.fit_wjet(pts, w, 3)
    nw = O.world_normals(n, Us, R)
    ang = O.unoriented_angle_deg(nw.numpy(), nw_ref.numpy())
    c = O.principal_curvatures(beta)
    ce = O.rectified_rel_err(c.numpy(), c_ref.numpy())
    print('%-7s %-7s  w maxdiff %.2e  R maxdiff %.2e  t2 %.2e | angle max %.2e p99 %.2e med %.2e | curv max %.2e med %.2e' % (mode_pt, mode_fc, (w-w_ref).abs().max(), (R-R_ref).abs().max(), (t2-t2_ref).abs().max(), ang.max(), np.quantile(ang,0.99), np.median(ang), ce.max(), np.median(ce)))
