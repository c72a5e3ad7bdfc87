"""Consumers of the JSON state log (api/helper/LogHelper.scala, PosteriorHelper.scala, the companion object of
sampling/loggers/JSONStateLogger.scala:205-236, helper/CallBackFunctions.scala): thinning of a logged chain, the logged
samples as shapes (model instances evaluated on the device), and the per-vertex posterior variance maps of
examples/DemoPosteriorVisualizationFemur.scala:17-28.  Host side except for Model.instance."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

from .io import JsonLogRecord, JSONStateLogger


def record_to_parameters(record: JsonLogRecord):
    """JSONStateLogger.jsonFormatToModelFittingParameters (:216-229).  A rejected record carries no parameters: the
    reference's `require` fails there, this raises ValueError."""
    from .api import ModelFittingParameters
    if len(record.rotation) != 3 or len(record.rotationCenter) != 3:
        raise ValueError("requirement failed: the log record has no pose (a rejected sample)")
    # the centre travels with the parameters; the device refuses a non-zero one (GiNGR always rotates about the origin,
    # GeneralRegistrationState.scala:147) instead of silently re-posing the shape
    return ModelFittingParameters(float(record.scaling), np.asarray(record.translation, dtype=np.float64),
                                  tuple(float(v) for v in record.rotation), np.asarray(record.modelParameters, dtype=np.float64),
                                  np.asarray(record.rotationCenter, dtype=np.float64))


def best_record(log: Sequence[JsonLogRecord]) -> JsonLogRecord:
    """JSONStateLogger.getBestStateFromLog (:231-234): sortBy(product).reverse.head -- the largest product; among equal
    values the LAST record of the log (stable sort, then reversed)."""
    if not log:
        raise ValueError("empty log")
    best = log[0]
    for r in log[1:]:
        if r.logvalue["product"] >= best.logvalue["product"]:
            best = r
    return best


def samples_from_log(log: Sequence[JsonLogRecord], takeEveryN: int = 50, total: int = 100, burnIn: int = 0
                     ) -> List[Tuple[JsonLogRecord, int]]:
    """LogHelper.samplesFromLog (:29-43): indices burnIn, burnIn + takeEveryN, ... below min(len, total), each moved back
    to the last ACCEPTED record at or before it (a rejected record holds no parameters; the chain's state there is the
    previous accepted one)."""
    out = []
    for i in range(burnIn, min(len(log), total), takeEveryN):
        j = i
        while not log[j].status:
            j -= 1
            if j < 0:
                raise IndexError("no accepted record at or before index %d" % i)
        out.append((log[j], j))
    return out[:min(total, len(out))]


def log_samples_to_shapes(model, records: Sequence[JsonLogRecord]) -> List[np.ndarray]:
    """LogHelper.logSamples2shapes (:45-53): modelInstanceShapePoseScale of every record (api.Model.instance, on the device)."""
    return [model.instance(record_to_parameters(r)) for r in records]


def vertex_normals(points, triangles) -> np.ndarray:
    """mesh.vertexNormals [scalismo-recalled, SURVEY A6]: normalised mean of the unit normals of the adjacent cells."""
    v = np.asarray(points, dtype=np.float64).reshape(-1, 3)
    t = np.asarray(triangles, dtype=np.int64).reshape(-1, 3)
    n = np.cross(v[t[:, 1]] - v[t[:, 0]], v[t[:, 2]] - v[t[:, 0]])
    n /= np.sqrt((n * n).sum(1))[:, None]
    acc = np.zeros_like(v)
    cnt = np.zeros(v.shape[0])
    for k in range(3):
        np.add.at(acc, t[:, k], n)
        np.add.at(cnt, t[:, k], 1.0)
    acc /= np.maximum(cnt, 1.0)[:, None]
    return acc / np.sqrt((acc * acc).sum(1))[:, None]


def distance_map_total(meshes: Sequence[np.ndarray]) -> np.ndarray:
    """PosteriorHelper.computeDistanceMapFromMeshesTotal (:31-53): per vertex the trace of the sample covariance
    (1 / (n - 1) normalisation) of its position over the meshes.  meshes: sequence of [M, 3] vertex arrays."""
    x = np.stack([np.asarray(m, dtype=np.float64).reshape(-1, 3) for m in meshes])         # [n, M, 3]
    n = x.shape[0]
    mean = x.sum(0) * (1.0 / n)
    d = x - mean
    return (d * d).sum(axis=(0, 2)) * (1.0 / (n - 1))


def distance_map_normal(meshes: Sequence[np.ndarray], triangles, reference: Optional[np.ndarray] = None,
                        sumNormals: bool = True) -> np.ndarray:
    """PosteriorHelper.computeDistanceMapFromMeshesNormal (:55-79): per vertex the sample variance of its position along a
    normal direction -- the mean of the samples' unit vertex normals (not re-normalised, as in the reference) when
    sumNormals, else the unit vertex normal of `reference`."""
    x = np.stack([np.asarray(m, dtype=np.float64).reshape(-1, 3) for m in meshes])
    n = x.shape[0]
    mean = x.sum(0) * (1.0 / n)
    if sumNormals:
        nrm = sum(vertex_normals(m, triangles) for m in x) * (1.0 / n)
    else:
        if reference is None:
            raise ValueError("sumNormals = False needs the reference mesh points")
        nrm = vertex_normals(reference, triangles)
    proj = ((x - mean) * nrm[None, :, :]).sum(2)
    return (proj * proj).sum(0) * (1.0 / (n - 1))


class SimpleLogger:
    """CallBackFunctions.SimpleLogger (:26-43): a chain-state callback that, every printUpdateFrequency states, prints the
    acceptance figures of the JSON logger, rewrites its file and evaluates the boundary-aware surface distance of the
    current fit against the target (comparison.RegistrationComparison; needs the meshes on the host)."""

    def __init__(self, jsonLogger: Optional[JSONStateLogger] = None, printUpdateFrequency: int = 100, comparison=None,
                 fit_triangles=None, target_mesh=None, out=print):
        self.jsonLogger, self.printUpdateFrequency = jsonLogger, int(printUpdateFrequency)
        self.comparison, self.fit_triangles, self.target_mesh, self.out = comparison, fit_triangles, target_mesh, out
        self.counter = 0

    def __call__(self, sample) -> None:
        self.counter += 1
        if self.counter % self.printUpdateFrequency == 0 and self.counter > 1:
            lg = self.jsonLogger
            if lg is not None:
                self.out(f"Total accepted ({lg.totalSamples}): {lg.percentAccepted}")
                for name in sorted(n for n in lg.generatedBy if n):
                    self.out(f"{name}: {lg.percentAcceptedOfType(name)}")
                if lg.path is not None:
                    lg.write()
            if self.comparison is not None and self.target_mesh is not None:
                avg, mx = self.comparison.evaluateReconstruction2GroundTruthBoundaryAware((sample.fit, self.fit_triangles),
                                                                                           self.target_mesh)
                self.out(f"average2surface: {avg} max: {mx}")
