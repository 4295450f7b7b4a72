"""Drop-in for the reference CLI ``tutorial/compute_normals.py``: same flags, same output files.

    python -m deepfit_b200.compute_normals --input cloud.xyz --output_path out/ --mode classic \
        --k_neighbors 256 --jet_order 3

writes ``<output_path>/<stem>.normals`` (N x 3) and ``<stem>.curv`` (N x 2) with ``np.savetxt``'s
default format, like reference lines :84-92.  Differences, all documented in DESIGN.md: BatchNorm
runs in eval mode; the QSTN rotation is cancelled on the normals (as test_n_est.py:154-157 and the
notebook do) unless ``--ref-compat`` reproduces the shipped script's omission; extra flags
``--precision`` and ``--batch``.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np
import torch


def build_parser():
    parser = argparse.ArgumentParser()
    parser.add_argument('--input', type=str, default='./0000000000.xyz', help='full path to input point cloud')
    parser.add_argument('--output_path', type=str, default='../log/outputs/', help='directory for the outputs')
    parser.add_argument('--gpu_idx', type=int, default=0, help='index of gpu to use (a GPU is required)')
    parser.add_argument('--trained_model_path', type=str, default='../trained_models/DeepFit',
                        help='path to trained model')
    parser.add_argument('--mode', type=str, default='classic', help='how to compute normals. use: DeepFit | classic')
    parser.add_argument('--k_neighbors', type=int, default=256, help='number of neighboring points for each query point')
    parser.add_argument('--jet_order', type=int, default=3,
                        help='order of jet to fit: 1-4. if in DeepFit mode, make sure to match training order')
    parser.add_argument('--compute_curvatures', type=bool, default=True, help='true | false indicator to compute curvatures')
    parser.add_argument('--ref-compat', action='store_true',
                        help='do not cancel the QSTN rotation on the normals (what the reference script does)')
    parser.add_argument('--precision', type=str, default='f16x3', help='f16x3 (fp32-grade, default) | bf16x3 | f16f8 | f16mix (faster, weights 1e-4 .. 4e-4 off) | bf16 (fastest, ~0.2 deg)')
    parser.add_argument('--batch', type=int, default=0, help='query points per chunk (0 = auto)')
    return parser


def main(argv=None):
    args = build_parser().parse_args(argv)
    if args.gpu_idx < 0:
        sys.exit("deepfit_b200 has no CPU path: --gpu_idx must be >= 0")
    from . import tutorial_utils as tu
    from .pipeline import NormalEstimator, model_from_state_dict

    device = torch.device("cuda:%d" % args.gpu_idx)
    torch.cuda.set_device(device)
    model = None
    if args.mode == 'DeepFit':
        params = torch.load(os.path.join(args.trained_model_path, 'DeepFit_params.pth'), weights_only=False)
        jet_order = params.jet_order
        print('Using {} order jet for surface fitting'.format(jet_order))
        checkpoint = torch.load(os.path.join(args.trained_model_path, 'DeepFit.pth'), map_location='cpu')
        if getattr(params, 'arch', 'simple') == '3dmfv' or params.point_tuple != 1:
            sys.exit("unsupported model configuration (arch=%s, point_tuple=%s)" % (params.arch, params.point_tuple))
        model = model_from_state_dict(checkpoint, args.k_neighbors, jet_order, params.weight_mode, args.precision, device)
        if not (params.points_per_patch == args.k_neighbors):
            print('Warning: You are using a different number of neighbors than trained.')
    else:
        jet_order = args.jet_order

    dataset = tu.SinglePointCloudDataset(args.input, points_per_patch=args.k_neighbors, device=device)
    est = NormalEstimator(dataset.points_gpu, args.k_neighbors, jet_order, args.mode, model, chunk=args.batch,
                          cancel_qstn=not args.ref_compat, grid=dataset.kdtree)
    if jet_order < 2:
        # the reference dies here with this very error (tutorial_utils.py:204-205) *before* writing anything
        raise ValueError("Can't compute curvatures for jet with order less than 2")
    normals, curvatures = est.run()
    normals = normals.cpu().numpy()
    curvatures = curvatures.cpu().numpy()

    os.makedirs(args.output_path, exist_ok=True)
    stem = os.path.splitext(os.path.basename(args.input))[0]
    normals_output_file_name = os.path.join(args.output_path, stem + '.normals')
    tu.savetxt(normals_output_file_name, normals)        # == np.savetxt(...), byte for byte
    print('Saved estimated normal vectors to ' + normals_output_file_name)
    curvatures_output_file_name = os.path.join(args.output_path, stem + '.curv')
    tu.savetxt(curvatures_output_file_name, curvatures)
    print('Saved estimated principal curvatures to ' + curvatures_output_file_name)


if __name__ == '__main__':
    main()
