"""coupe.dvsg_b200 -- B200-native (sm_100a) frame-warping hot path of DVSG.

Drop-in modules with the reference's names and signatures:
    coupe.dvsg_b200.ThinPlateSpline.ThinPlateSpline(U, coord, vector, out_size)
    coupe.dvsg_b200.ThinPlateSpline2.ThinPlateSpline2(U, source, target, out_size)
    coupe.dvsg_b200.spatial_transformer._meshgrid / _interpolate / bilinear_interp /
                                         ProjectiveTransformer / AffineTransformer
    coupe.dvsg_b200.warp_with_optical_flow.tf_warp(im, flow, out_height, out_width)
All of them call hand-written CUDA kernels through the C ABI of include/dvsg_warp.h.
"""
__version__ = '0.1.0'
