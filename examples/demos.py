"""Counterparts of the reference's demos (examples/DemoCPD.scala, DemoICP.scala, DemoMultiResolution.scala,
DemoLandmarks.scala) on the device library, without the UI: same data sets, kernels, configurations and run schedule,
results printed instead of shown.  Needs a B200 and the built library (python -m gingr_b200.build).

    python examples/demos.py cpd      --data /path/to/GiNGR/examples/data        # femur, Gauss(50, 70), 100 / 100 points
    python examples/demos.py icp      --data ...  [--log chain.json]
    python examples/demos.py multires --data ...                                   # bunny: CPD 100 -> CPD 500 -> ICP 1000
    python examples/demos.py landmarks --data ...                                  # femur with / without landmarks
    python examples/demos.py posterior --data ... --log chain.json                 # DemoPosteriorVisualizationFemur: variance maps
    python examples/demos.py gpmm --data ... --cache models/                       # Create{Femur, Bunny}GPMM: build + cache the models
    python examples/demos.py cpd                                                   # no --data: a synthetic sphere pair

Substitutions: meshes are decimated by gingr_b200.decimate instead of scalismo's quadric decimation; the
bunny's target.ply is not shipped with the reference, so the bunny target is the bunny itself under a smooth synthetic
deformation plus the demo's rigid offset."""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

STATUS = {0: "None", 1: "MaxIteration", 2: "Converged", 3: "ModelFlexibilityError"}


def load_dataset(name: str, data_dir, offset=None):
    """DemoDatasetLoader.{femur, bunny} (examples/DemoHelper/DemoDatasetLoader.scala:107-160) -> dict with the reference
    and target meshes, their landmarks (or None) and the default Gaussian kernel of the data set."""
    from gingr_b200 import io, synthetic
    rot, trans = offset if offset is not None else (np.eye(3), np.zeros(3))
    if data_dir is None or name == "synthetic":
        rv, rt = synthetic.sphere_mesh(2000)
        tv, tt = synthetic.sphere_mesh(2400)
        tv = synthetic.make_target(tv, 0)
        return dict(name="synthetic", ref=(rv, rt), target=(tv @ rot.T + trans, tt), ref_lms=None, target_lms=None,
                    kernel=(50.0, 70.0))
    path = os.path.join(data_dir, name)
    if name == "femur":
        rv, rt = io.read_mesh(os.path.join(path, "femur.stl"))
        tv, tt = io.read_mesh(os.path.join(path, "femur_target.stl"))
        rl = io.read_landmarks(os.path.join(path, "femur.json"))
        tl = [io.Landmark(l.id, rot @ l.point + trans, l.covariance) for l in io.read_landmarks(os.path.join(path, "femur_target.json"))]
        return dict(name=name, ref=(rv, rt), target=(tv @ rot.T + trans, tt), ref_lms=rl, target_lms=tl, kernel=(50.0, 70.0))
    if name == "bunny":
        rv, rt = io.read_mesh(os.path.join(path, "bunny.ply"))
        tgt = os.path.join(path, "target.ply")
        if os.path.exists(tgt):
            tv, tt = io.read_mesh(tgt)
        else:
            rng = np.random.default_rng(0)
            tv, tt = rv + synthetic.smooth_displacement(rv, rng), rt
        return dict(name=name, ref=(rv, rt), target=(tv @ rot.T + trans, tt), ref_lms=None, target_lms=None, kernel=(20.0, 40.0))
    raise SystemExit(f"unknown data set {name}")


def build(ctx, ds, cache_dir=None):
    """(model, target): modelGauss() of the data set -- from the cache file when there is one -- and the uploaded target."""
    from gingr_b200 import api, io
    scaling, sigma = ds["kernel"]
    t0 = time.time()
    if cache_dir is not None:
        model = io.load_or_create_model(ctx, cache_dir, ds["name"], ds["ref"][0], ds["ref"][1], api.GaussKernel(scaling, sigma))
    else:
        model = api.SimpleTriangleModels3D.create(ctx, ds["ref"][0], ds["ref"][1], api.GaussKernel(scaling, sigma))
    print(f"model: {model.M} points, rank {model.rank} (Gauss scaling {scaling}, sigma {sigma}) in {time.time() - t0:.2f} s")
    return model, api.Target(ctx, ds["target"][0], ds["target"][1])


def print_status(tag, st, t0, ctx=None, model=None, ds=None):
    """general.printStatus() plus the distances SimpleRegistrator.run prints (RegistrationComparison ...BoundaryAware)."""
    line = (f"{tag}: status {STATUS.get(st.status, st.status)}, iteration {st.iteration}, sigma2 {st.sigma2:.6g}, "
            f"{time.time() - t0:.2f} s")
    if ctx is not None and model is not None and model.triangles is not None:
        from gingr_b200.comparison import RegistrationComparison
        avg, mx = RegistrationComparison(ctx).evaluateReconstruction2GroundTruthBoundaryAware((st.fit, model.triangles), ds["target"])
        line += f", average2surface {avg:.4f} max {mx:.4f}"
    print(line)


def demo_cpd_or_icp(args, which):
    from gingr_b200 import api
    ctx = api.Context(0)
    ds = load_dataset("femur" if args.data else "synthetic", args.data)
    model, target = build(ctx, ds, args.cache)
    gi = api.GingrInterface(ctx, model, target, evaluatorUncertainty=5.0, logFileFittingParameters=args.log)
    if which == "cpd":
        det, pro = api.CpdConfiguration(maxIterations=100, initialSigma=1.0), api.CpdConfiguration(maxIterations=args.samples, initialSigma=1.0)
        make = gi.CPD
    else:
        det = api.IcpConfiguration(maxIterations=100, initialSigma=1.0, endSigma=1.0)
        pro = api.IcpConfiguration(maxIterations=args.samples, initialSigma=1.0, endSigma=1.0)
        make = gi.ICP
    t0 = time.time()
    best = make(det).runDecimated(100, 100, globalTransformation=api.NO_TRANSFORMS)
    print_status("deterministic", best, t0, ctx, model, ds)
    t0 = time.time()
    sr = make(pro)
    best = sr.runDecimated(100, 100, globalTransformation=api.NO_TRANSFORMS, probabilistic=True, seed=args.seed)
    print_status("probabilistic (best sample)", best, t0, ctx, model, ds)
    if sr.jsonLogger is not None:
        print(f"accepted {sr.jsonLogger.percentAccepted:.2f} of {sr.jsonLogger.totalSamples} samples, log written to {args.log}")


def demo_multires(args):
    from gingr_b200 import api, rotation
    ctx = api.Context(0)
    offset = (rotation.euler_to_matrix(0.1, 0.1, 0.1), np.array([50.0, 50.0, 50.0]))       # DemoMultiResolution.scala:16
    ds = load_dataset("bunny" if args.data else "synthetic", args.data, offset)
    model, target = build(ctx, ds, args.cache)
    gi = api.GingrInterface(ctx, model, target)
    t0 = time.time()
    coarse = gi.CPD(api.CpdConfiguration(maxIterations=50)).runDecimated(100, 100, globalTransformation=api.RIGID_TRANSFORMS)
    print_status("coarse (CPD 100)", coarse, t0, ctx, model, ds)
    t0 = time.time()
    medium = gi.CPD(api.CpdConfiguration(maxIterations=50, initialSigma=coarse.sigma2)).runDecimated(
        500, 500, generalState=coarse, globalTransformation=api.RIGID_TRANSFORMS)
    print_status("medium (CPD 500)", medium, t0, ctx, model, ds)
    t0 = time.time()
    fine = gi.ICP(api.IcpConfiguration(maxIterations=100, initialSigma=2.0, endSigma=0.01)).runDecimated(
        1000, 1000, generalState=medium, globalTransformation=api.NO_TRANSFORMS)   # argument ignored as in the reference: the
    # handed-over state keeps its RigidTransforms (SimpleRegistrator.scala:93-95)
    print_status("fine (ICP 1000)", fine, t0, ctx, model, ds)


def demo_landmarks(args):
    from gingr_b200 import api
    ctx = api.Context(0)
    ds = load_dataset("femur" if args.data else "synthetic", args.data)
    model, target = build(ctx, ds, args.cache)
    cfg = api.CpdConfiguration(maxIterations=30)
    t0 = time.time()
    res = api.GingrInterface(ctx, model, target, evaluatorUncertainty=2.0).CPD(cfg).runDecimated(100, 100)
    print_status("without landmarks", res, t0, ctx, model, ds)
    if ds["ref_lms"] is None:
        print("this data set has no landmarks")
        return
    t0 = time.time()
    res = api.GingrInterface(ctx, model, target, modelLandmarks=ds["ref_lms"], targetLandmarks=ds["target_lms"],
                             evaluatorUncertainty=2.0).CPD(cfg).runDecimated(100, 100)
    print_status("with landmarks", res, t0, ctx, model, ds)


def demo_gpmm(args):
    """examples/Create{Femur, Bunny}GPMM.scala: build the data set's Gaussian-kernel model on the device and leave it in the
    cache directory (the file DataSetLoader.model looks for), then report its leading variances."""
    from gingr_b200 import api
    ctx = api.Context(0)
    for name in (("femur", "bunny") if args.data else ("synthetic",)):
        ds = load_dataset(name, args.data)
        model, target = build(ctx, ds, args.cache or ".")
        _, _, _, var = model.download()
        print(f"{ds['name']}: leading standard deviations {np.sqrt(var[:5]).round(3).tolist()}, last {float(np.sqrt(var[-1])):.4g}")
        target.close()
        model.close()


def demo_posterior(args):
    """DemoICP's probabilistic run with its JSON log, then examples/DemoPosteriorVisualizationFemur.scala:10-28: thin the
    chain after a burn-in, turn the samples into shapes and reduce them to per-vertex variance maps (no UI: the maps are
    written as .npy next to the log)."""
    from gingr_b200 import api, helper, io
    ctx = api.Context(0)
    ds = load_dataset("femur" if args.data else "synthetic", args.data)
    model, target = build(ctx, ds, args.cache)
    log_path = args.log or "targetFittingICP.json"
    gi = api.GingrInterface(ctx, model, target, evaluatorUncertainty=5.0, logFileFittingParameters=log_path)
    cfg = api.IcpConfiguration(maxIterations=args.samples, initialSigma=1.0, endSigma=1.0)
    t0 = time.time()
    best = gi.ICP(cfg).runDecimated(100, 100, globalTransformation=api.NO_TRANSFORMS, probabilistic=True, seed=args.seed)
    print_status("probabilistic (best sample)", best, t0, ctx, model, ds)
    full_log = io.JSONStateLogger.load(log_path)
    burn_in = min(100, len(full_log) // 4)
    samples = helper.samples_from_log(full_log, takeEveryN=max(1, len(full_log) // 40), total=10000, burnIn=burn_in)
    print(f"Number of samples from log: {len(samples)}/{len(full_log) - burn_in}")
    shapes = helper.log_samples_to_shapes(model, [r for r, _ in samples])
    best_pars = helper.record_to_parameters(helper.best_record([r for r in full_log if r.status]))
    best_shape = model.instance(best_pars)
    total = helper.distance_map_total(shapes)
    normal = helper.distance_map_normal(shapes, model.triangles) if model.triangles is not None else total
    base = os.path.splitext(log_path)[0]
    np.save(base + "_variance_total.npy", total)
    np.save(base + "_variance_normal.npy", normal)
    np.save(base + "_best_shape.npy", best_shape)
    print(f"posterior variance per vertex: total mean {total.mean():.4g} max {total.max():.4g}; along normals mean "
          f"{normal.mean():.4g} max {normal.max():.4g}; written to {base}_variance_*.npy")


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("demo", choices=["cpd", "icp", "multires", "landmarks", "posterior", "gpmm"])
    ap.add_argument("--data", default=None, help="the reference's examples/data directory (default: synthetic spheres)")
    ap.add_argument("--cache", default=None, help="directory for the cached model files (DemoDatasetLoader.model)")
    ap.add_argument("--log", default=None, help="JSON state log of the probabilistic run (logFileFittingParameters)")
    ap.add_argument("--samples", type=int, default=1000, help="maxIterations of the probabilistic run")
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args(argv)
    if args.demo in ("cpd", "icp"):
        demo_cpd_or_icp(args, args.demo)
    elif args.demo == "multires":
        demo_multires(args)
    elif args.demo == "posterior":
        demo_posterior(args)
    elif args.demo == "gpmm":
        demo_gpmm(args)
    else:
        demo_landmarks(args)


if __name__ == "__main__":
    main()
