"""Generate tests/golden/*.npz by executing the UNMODIFIED reference sources.

Runs only in the build container (needs /root/reference).  TensorFlow is absent there,
so the reference files are executed against oracle/tf1_shim.py (a TF1-API emulation on
torch CPU, fp32) -- see that file's header for what this does and does not pin.
Gradients are torch autograd over the reference's own op sequence, i.e. what
`tf.gradients` would assemble from the same graph.

    python tests/golden/make_golden.py          # rewrites tests/golden/*.npz
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import tf1_shim as tf  # noqa: E402

REF = os.environ.get('DVSG_REFERENCE', '/root/reference')
OUT = os.path.dirname(os.path.abspath(__file__))


def smooth_image(rng, b, h, w, c):
    """Band-limited frames in [0,1]: low-frequency sinusoids + 1% noise."""
    yy, xx = np.meshgrid(np.arange(h), np.arange(w), indexing='ij')
    im = np.zeros((b, h, w, c), np.float64)
    for bi in range(b):
        for ci in range(c):
            for _ in range(4):
                fx, fy = rng.uniform(-1, 1, 2) / 12.0
                im[bi, :, :, ci] += np.sin(2 * np.pi * (fx * xx + fy * yy) + rng.uniform(0, 6.28))
    im = (im - im.min()) / (im.max() - im.min())
    return (0.99 * im + 0.01 * rng.random(im.shape)).astype(np.float32)


def mesh(n_rows, n_cols, b):
    xs = np.linspace(-1, 1, n_cols)
    ys = np.linspace(-1, 1, n_rows)
    gx, gy = np.meshgrid(xs, ys)
    m = np.stack([gx.reshape(-1), gy.reshape(-1)], 1).astype(np.float32)
    return np.tile(m[None], (b, 1, 1))


def tt(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def run_tps(fn, u, coord, second, out_size, g_out, g_x, g_y, dtype):
    tf.set_float(dtype)
    u_t = tt(u).to(dtype).requires_grad_(True)
    s_t = tt(second).to(dtype).requires_grad_(True)
    c_t = tt(coord).to(dtype)
    out, x, y = fn(u_t, c_t, s_t, list(out_size))
    loss = (out * tt(g_out).to(dtype)).sum() + (x * tt(g_x).to(dtype)).sum() + (y * tt(g_y).to(dtype)).sum()
    loss.backward()
    tf.set_float(torch.float32)
    return [v.detach().numpy() for v in (out, x, y, u_t.grad, s_t.grad)]


def tps_cases(rng):
    ref_tps = tf.load_reference(os.path.join(REF, 'ThinPlateSpline.py'), 'ref_ThinPlateSpline')
    ref_tps2 = tf.load_reference(os.path.join(REF, 'ThinPlateSpline2.py'), 'ref_ThinPlateSpline2')
    cases = {}
    specs = [
        # name,           B, H,  W,  C, mesh,   out_size, amplitude, variant
        ('tps_4x4',       2, 24, 32, 3, (4, 4), (24, 32), 0.10, 1),
        ('tps_5x5',       2, 18, 20, 3, (5, 5), (18, 20), 0.10, 1),
        ('tps_4x4_big',   1, 36, 52, 3, (4, 4), (36, 52), 0.35, 1),   # samples far outside the frame
        ('tps_resize',    2, 16, 20, 2, (3, 3), (12, 10), 0.05, 1),   # out_size != in size, C=2
        ('tps2_4x4',      2, 20, 28, 3, (4, 4), (20, 28), 0.10, 2),
        ('tps_8x8',       1, 20, 24, 1, (8, 8), (20, 24), 0.03, 1),   # N=67 > one warp
    ]
    for name, b, h, w, c, (mr, mc), osz, amp, variant in specs:
        u = smooth_image(rng, b, h, w, c)
        coord = mesh(mr, mc, b)
        vec = rng.uniform(-amp, amp, coord.shape).astype(np.float32)
        second = vec if variant == 1 else (coord + vec).astype(np.float32)
        fn = ref_tps.ThinPlateSpline if variant == 1 else ref_tps2.ThinPlateSpline2
        n_out = b * osz[0] * osz[1]
        g_out = rng.standard_normal((b, osz[0], osz[1], c)).astype(np.float32)
        g_x = (rng.standard_normal(n_out) * 0.1).astype(np.float32)
        g_y = (rng.standard_normal(n_out) * 0.1).astype(np.float32)
        out, x, y, gu, gs = run_tps(fn, u, coord, second, osz, g_out, g_x, g_y, torch.float32)
        out64, x64, y64, gu64, gs64 = run_tps(fn, u, coord, second, osz, g_out, g_x, g_y, torch.float64)
        cases[name] = dict(u=u, coord=coord, second=second, out_size=np.array(osz), variant=np.array(variant),
                           g_out=g_out, g_x=g_x, g_y=g_y, out=out, x=x, y=y, grad_u=gu, grad_second=gs,
                           out64=out64, x64=x64, y64=y64, grad_u64=gu64, grad_second64=gs64)
    # validity mask: all-ones image (model.py:82,85)
    b, h, w = 1, 24, 32
    coord = mesh(4, 4, b)
    vec = rng.uniform(-0.15, 0.15, coord.shape).astype(np.float32)
    ones = np.ones((b, h, w, 3), np.float32)
    out, x, y = ref_tps.ThinPlateSpline(tt(ones), tt(coord), tt(vec), [h, w])
    cases['tps_mask'] = dict(u=ones, coord=coord, second=vec, out_size=np.array([h, w]), variant=np.array(1),
                             out=out.numpy(), x=x.numpy(), y=y.numpy())
    return cases


def sampler_cases(rng):
    st = tf.load_reference(os.path.join(REF, 'spatial_transformer.py'), 'ref_spatial_transformer')
    fl = tf.load_reference(os.path.join(REF, 'warp_with_optical_flow.py'), 'ref_warp_with_optical_flow')
    cases = {}
    # B1
    cases['meshgrid'] = dict(out_size=np.array([7, 9]), grid=st._meshgrid([7, 9]).numpy(),
                             out_size2=np.array([288, 512]), grid2=st._meshgrid([288, 512]).numpy()[::97].copy())
    # B2 / B3: C = 18 is the only live call (model.py:160-164); coordinates reach far outside
    for name, b, h, w, c, osz in [('bilinear_c18', 2, 14, 18, 18, (14, 18)), ('bilinear_c3', 2, 15, 21, 3, (9, 13))]:
        im = smooth_image(rng, b, h, w, c)
        n = b * osz[0] * osz[1]
        x = rng.uniform(-1.3, 1.3, n).astype(np.float32)
        y = rng.uniform(-1.3, 1.3, n).astype(np.float32)
        x[:8] = [-1.0, 1.0, -1.0 - 2.0 / (w - 1), 1.0 + 2.0 / (w - 1), 0.0, -5.0, 5.0, 1.0]
        y[:8] = [-1.0, 1.0, 0.0, 0.0, 1.0 + 2.0 / (h - 1), 0.3, -0.3, -1.0]
        g = rng.standard_normal((n, c)).astype(np.float32)
        im_t, x_t, y_t = tt(im).requires_grad_(True), tt(x).requires_grad_(True), tt(y).requires_grad_(True)
        out = st._interpolate(im_t, x_t, y_t, list(osz), 'bilinear')
        (out * tt(g)).sum().backward()
        cases[name] = dict(im=im, x=x, y=y, out_size=np.array(osz), g_out=g, out=out.detach().numpy(),
                           grad_im=im_t.grad.numpy(), grad_x=x_t.grad.numpy(), grad_y=y_t.grad.numpy())
    # N2: projective / affine transformers on top of B2
    b, h, w, c = 2, 12, 16, 18
    im = smooth_image(rng, b, h, w, c)
    theta = (rng.uniform(-1, 1, (b, 8)) * np.array([0.1, 0.1, 0.5, 0.1, 0.1, 0.5, 0.1, 0.1])
             + np.array([1.0, 0, 0, 0, 1.0, 0, 0, 0])).astype(np.float32)          # model.py:161-163
    pt = st.ProjectiveTransformer([h, w])
    cases['projective'] = dict(im=im, theta=theta, out_size=np.array([h, w]),
                               out=pt.transform(tt(im), tt(theta)).numpy())
    theta6 = (rng.uniform(-0.2, 0.2, (b, 6)) + np.array([1.0, 0, 0, 0, 1.0, 0])).astype(np.float32)
    at = st.AffineTransformer([h, w])
    cases['affine'] = dict(im=im, theta=theta6, out_size=np.array([h, w]),
                           out=at.transform(tt(im), tt(theta6)).numpy())
    # C1
    for name, b, h, w, c, amp in [('flow_small', 2, 16, 24, 3, 2.0), ('flow_large', 1, 18, 22, 3, 12.0)]:
        im = smooth_image(rng, b, h, w, c)
        flow = rng.uniform(-amp, amp, (b, h, w, 2)).astype(np.float32)
        flow[0, 0, 0] = [-1.0, -1.0]
        flow[0, 0, 1] = [0.0, 0.0]
        flow[0, 1, 0] = [w + 3.0, 0.5]
        g = rng.standard_normal((b, h, w, c)).astype(np.float32)
        im_t, f_t = tt(im).requires_grad_(True), tt(flow).requires_grad_(True)
        out = fl.tf_warp(im_t, f_t, h, w)
        (out * tt(g)).sum().backward()
        cases[name] = dict(im=im, flow=flow, g_out=g, out=out.detach().numpy(),
                           grad_im=im_t.grad.numpy(), grad_flow=f_t.grad.numpy())
    return cases


def loss_cases():
    """N3: Trainer.masked_MSE / temporal_loss / get_surf_loss, trainer.py:232-250,363-386 --
    the method bodies executed UNMODIFIED (cut out of trainer.py's AST: the file itself imports tensorlayer).
    Own generator, so that adding these cases leaves the older fixtures bit-identical."""
    import types
    rng = np.random.default_rng(20261019)
    fl = tf.load_reference(os.path.join(REF, 'warp_with_optical_flow.py'), 'ref_warp_with_optical_flow')
    fns = tf.load_reference_methods(os.path.join(REF, 'trainer.py'), 'Trainer', ['masked_MSE', 'temporal_loss', 'get_surf_loss'],
                                    {'tf_warp': fl.tf_warp})
    me = types.SimpleNamespace()
    for k, f in fns.items():
        setattr(me, k, types.MethodType(f, me))
    cases = {}
    # masked_MSE: fractional validity mask (a warp of ones), one frame fully masked out (div_no_nan)
    b, h, w, c = 3, 20, 28, 3
    pred, gt = smooth_image(rng, b, h, w, c), smooth_image(rng, b, h, w, c)
    mask = np.clip(smooth_image(rng, b, h, w, c) * 1.6 - 0.3, 0.0, 1.0).astype(np.float32)
    mask[1] = 0.0
    p_t, g_t, m_t = (tt(a).requires_grad_(True) for a in (pred, gt, mask))
    loss = me.masked_MSE(p_t, g_t, m_t, 'loss')
    (loss * 1.5).backward()
    cases['loss_masked_mse'] = dict(pred=pred, gt=gt, mask=mask, loss=loss.detach().numpy(), grad_scale=np.float32(1.5),
                                    grad_pred=p_t.grad.numpy(), grad_gt=g_t.grad.numpy(), grad_mask=m_t.grad.numpy())
    # temporal_loss: tf_warp of the prediction and of its mask, then the masked MSE
    b, h, w, c = 2, 18, 24, 3
    pred, gt = smooth_image(rng, b, h, w, c), smooth_image(rng, b, h, w, c)
    mask_pred = np.clip(smooth_image(rng, b, h, w, c) * 1.5 - 0.2, 0.0, 1.0).astype(np.float32)
    mask_gt = (rng.random((b, h, w, c)) > 0.25).astype(np.float32)
    flow = rng.uniform(-2.5, 2.5, (b, h, w, 2)).astype(np.float32)
    p_t, mp_t = tt(pred).requires_grad_(True), tt(mask_pred).requires_grad_(True)
    loss = me.temporal_loss(p_t, tt(gt), mp_t, tt(mask_gt), tt(flow), h, w, 'loss')
    loss.backward()
    cases['loss_temporal'] = dict(pred=pred, gt=gt, mask_pred=mask_pred, mask_gt=mask_gt, flow=flow, loss=loss.detach().numpy(),
                                  grad_pred=p_t.grad.numpy(), grad_mask_pred=mp_t.grad.numpy())
    # get_surf_loss: dense grids x, y (flat [B*h*w]), feature lists padded with the sentinel index h*w
    b, h, w, p, n_pad = 3, 16, 22, 40, 6
    x = (np.tile(np.linspace(-1, 1, w), (b, h, 1)) + rng.uniform(-0.05, 0.05, (b, h, w))).astype(np.float32).reshape(-1)
    y = (np.tile(np.linspace(-1, 1, h)[:, None], (b, 1, w)) + rng.uniform(-0.05, 0.05, (b, h, w))).astype(np.float32).reshape(-1)
    surf = np.zeros((b, 2, p, 2), np.int32)
    surf[:, :, :, 0] = rng.integers(0, w, (b, 2, p))
    surf[:, :, :, 1] = rng.integers(0, h, (b, 2, p))
    surf[:, 0, -n_pad:, :] = 0                  # data_loader.py:295-308 pads: unstable (0,0), stable index h*w
    surf[:, 1, -n_pad:, 0] = 0
    surf[:, 1, -n_pad:, 1] = h
    max_dim = np.array([p - n_pad, 0, p], np.float32)
    x_t, y_t = tt(x).requires_grad_(True), tt(y).requires_grad_(True)
    loss = me.get_surf_loss(tt(surf), x_t, y_t, tt(max_dim), b, w, h)
    loss.backward()
    cases['loss_surf'] = dict(surf=surf, x=x, y=y, max_dim=max_dim, hw=np.array([h, w]), loss=loss.detach().numpy(),
                              grad_x=x_t.grad.numpy(), grad_y=y_t.grad.numpy())
    return cases


def elastic_cases():
    """N4: ElasticTransformer (spatial_transformer.py:93-362) executed unmodified, with autograd
    gradients w.r.t. the input and theta.  Own generator: the older fixtures stay bit-identical."""
    rng = np.random.default_rng(20261020)
    st = tf.load_reference(os.path.join(REF, 'spatial_transformer.py'), 'ref_spatial_transformer_elastic')
    cases = {}
    for name, b, h, w, c, g, osz, amp in [('elastic_4x4', 2, 14, 18, 3, 4, (14, 18), 0.12), ('elastic_3x3_resize', 1, 12, 16, 18, 3, (10, 12), 0.2)]:
        im = smooth_image(rng, b, h, w, c)
        theta = rng.uniform(-amp, amp, (b, 2 * g * g)).astype(np.float32)
        g_out = rng.standard_normal((b, osz[0], osz[1], c)).astype(np.float32)
        et = st.ElasticTransformer(list(osz), param_dim=2 * g * g, param_dim_per_side=g)
        im_t, th_t = tt(im).requires_grad_(True), tt(theta).requires_grad_(True)
        out, x_s, y_s = et.transform(im_t, th_t)
        (out * tt(g_out)).sum().backward()
        cases[name] = dict(im=im, theta=theta, grid_size=np.array(g), out_size=np.array(osz), g_out=g_out, out=out.detach().numpy(),
                           x=x_s.detach().numpy(), y=y_s.detach().numpy(), grad_im=im_t.grad.numpy(), grad_theta=th_t.grad.numpy(),
                           l_inv=et.L_inv.detach().numpy(), abs_theta=et.get_abs_theta(tt(theta)).numpy(),
                           abs_src=et.get_abs_src_points(b).numpy())
    return cases


def transformer_grad_cases():
    """Backward of ProjectiveTransformer / AffineTransformer (spatial_transformer.py:5-91, 364-452) w.r.t. the input and
    theta: the reference's transform() executed unmodified, gradients by autograd over its op sequence (matmul, div_no_nan,
    bilinear_interp).  No reference caller differentiates them (model.py:156-167 feeds random constants); built for
    completeness.  Own generator: the older fixtures stay bit-identical.  `grad_theta_grid` isolates the grid stage
    (_transform alone, random upstream gradients on x_s and y_s)."""
    rng = np.random.default_rng(20261021)
    st = tf.load_reference(os.path.join(REF, 'spatial_transformer.py'), 'ref_spatial_transformer_grad')
    cases = {}
    for name, proj, b, h, w, c, osz in [('projective_grad', True, 3, 12, 16, 3, (12, 16)), ('projective_grad_c18', True, 2, 10, 14, 18, (9, 11)),
                                        ('affine_grad', False, 2, 12, 16, 3, (10, 18))]:
        im = smooth_image(rng, b, h, w, c)
        if proj:
            theta = (rng.uniform(-1, 1, (b, 8)) * np.array([0.1, 0.1, 0.5, 0.1, 0.1, 0.5, 0.1, 0.1])
                     + np.array([1.0, 0, 0, 0, 1.0, 0, 0, 0])).astype(np.float32)      # model.py:161-163
            tr = st.ProjectiveTransformer(list(osz))
        else:
            theta = (rng.uniform(-0.2, 0.2, (b, 6)) + np.array([1.0, 0, 0, 0, 1.0, 0])).astype(np.float32)
            tr = st.AffineTransformer(list(osz))
        g_out = rng.standard_normal((b, osz[0], osz[1], c)).astype(np.float32)
        im_t, th_t = tt(im).requires_grad_(True), tt(theta).requires_grad_(True)
        out = tr.transform(im_t, th_t)
        (out * tt(g_out)).sum().backward()
        x_s, y_s = tr._transform(tt(im), tt(theta))
        # the grid stage alone: gradient of sum(x_s * g_x + y_s * g_y) w.r.t. theta
        g_x = rng.standard_normal(x_s.shape).astype(np.float32)
        g_y = rng.standard_normal(y_s.shape).astype(np.float32)
        th2 = tt(theta).requires_grad_(True)
        x2, y2 = tr._transform(tt(im), th2)
        ((x2 * tt(g_x)).sum() + (y2 * tt(g_y)).sum()).backward()
        cases[name] = dict(im=im, theta=theta, out_size=np.array(osz), g_out=g_out, out=out.detach().numpy(), x=x_s.numpy(), y=y_s.numpy(),
                           grad_im=im_t.grad.numpy(), grad_theta=th_t.grad.numpy(), g_x=g_x, g_y=g_y, grad_theta_grid=th2.grad.numpy())
    return cases


def coord_grad_cases():
    """Gradient w.r.t. the control-point positions `coord` of ThinPlateSpline / ThinPlateSpline2 (no reference caller takes
    it: model.py:62-68 builds the mesh as a constant): the reference executed unmodified with coord.requires_grad, fp32 and
    fp64.  Irregular meshes (a regular mesh + jitter), one per frame.  Own generator: the older fixtures stay bit-identical."""
    rng = np.random.default_rng(20261022)
    ref_tps = tf.load_reference(os.path.join(REF, 'ThinPlateSpline.py'), 'ref_ThinPlateSpline_cg')
    ref_tps2 = tf.load_reference(os.path.join(REF, 'ThinPlateSpline2.py'), 'ref_ThinPlateSpline2_cg')
    cases = {}
    for name, b, h, w, c, (mr, mc), osz, amp, variant in [('tps_coord_grad', 2, 20, 28, 3, (4, 4), (20, 28), 0.10, 1),
                                                           ('tps2_coord_grad', 2, 18, 24, 3, (3, 4), (16, 20), 0.08, 2)]:
        u = smooth_image(rng, b, h, w, c)
        coord = (mesh(mr, mc, b) + rng.uniform(-0.06, 0.06, (b, mr * mc, 2))).astype(np.float32)
        vec = rng.uniform(-amp, amp, coord.shape).astype(np.float32)
        second = vec if variant == 1 else (coord + vec).astype(np.float32)
        fn = ref_tps.ThinPlateSpline if variant == 1 else ref_tps2.ThinPlateSpline2
        n_out = b * osz[0] * osz[1]
        g_out = rng.standard_normal((b, osz[0], osz[1], c)).astype(np.float32)
        g_x = (rng.standard_normal(n_out) * 0.1).astype(np.float32)
        g_y = (rng.standard_normal(n_out) * 0.1).astype(np.float32)
        res = {}
        for tag, dtype in (('', torch.float32), ('64', torch.float64)):
            tf.set_float(dtype)
            u_t = tt(u).to(dtype).requires_grad_(True)
            s_t = tt(second).to(dtype).requires_grad_(True)
            c_t = tt(coord).to(dtype).requires_grad_(True)
            out, x, y = fn(u_t, c_t, s_t, list(osz))
            ((out * tt(g_out).to(dtype)).sum() + (x * tt(g_x).to(dtype)).sum() + (y * tt(g_y).to(dtype)).sum()).backward()
            tf.set_float(torch.float32)
            res.update({'out' + tag: out.detach().numpy(), 'x' + tag: x.detach().numpy(), 'y' + tag: y.detach().numpy(),
                        'grad_u' + tag: u_t.grad.numpy(), 'grad_second' + tag: s_t.grad.numpy(), 'grad_coord' + tag: c_t.grad.numpy()})
        cases[name] = dict(u=u, coord=coord, second=second, out_size=np.array(osz), variant=np.array(variant), g_out=g_out, g_x=g_x, g_y=g_y, **res)
    return cases


def main():
    rng = np.random.default_rng(20261018)
    allc = {}
    allc.update(tps_cases(rng))
    allc.update(sampler_cases(rng))
    allc.update(loss_cases())
    allc.update(elastic_cases())
    allc.update(transformer_grad_cases())
    allc.update(coord_grad_cases())
    for name, arrays in allc.items():
        path = os.path.join(OUT, name + '.npz')
        np.savez_compressed(path, **arrays)
        print('%-16s %8.1f KB' % (name, os.path.getsize(path) / 1024.0))


if __name__ == '__main__':
    main()
