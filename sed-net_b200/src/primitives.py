"""Point -> primitive distances with the reference's class and method names (reference src/primitives.py:18-206),
evaluated by the elementwise kernel of libsednet_b200.so."""
import torch

from . import _lib

_PLANE, _CONE, _CYLINDER, _SPHERE, _TORUS = 1, 3, 4, 5, 7


def _flat(params):
    return torch.cat([torch.as_tensor(p, dtype=torch.float32).reshape(-1).to(_dev(params)) for p in params])


def _dev(params):
    for p in params:
        if isinstance(p, torch.Tensor) and p.is_cuda:
            return p.device
    return torch.device("cuda", torch.cuda.current_device())


class ComputePrimitiveDistance:
    def __init__(self, reduce=True, one_side=False):
        self.reduce = reduce
        self.one_side = one_side

    def _run(self, prim, points, flat, sqrt):
        points = _lib.require_cuda(points, name="points")
        q = torch.zeros(8, dtype=torch.float32, device=points.device)
        q[: flat.numel()] = flat.to(points.device)
        out = torch.empty(points.shape[0], dtype=torch.float32, device=points.device)
        _lib.call("sed_primitive_distance", _lib.ptr(points), points.shape[0], prim, _lib.ptr(q), int(bool(sqrt)),
                  _lib.ptr(out), _lib.stream())
        return torch.mean(out) if self.reduce else out

    def distance_from_torus(self, points, params, sqrt=False):
        """src/primitives.py:58-87: params = (axis, center, major_radius, minor_radius)."""
        return self._run(_TORUS, points, _flat(params), sqrt)

    def distance_from_plane(self, points, params, sqrt=False):
        """src/primitives.py:89-111: params = (a, d)."""
        return self._run(_PLANE, points, _flat(params), sqrt)

    def distance_from_sphere(self, points, params, sqrt=False):
        """src/primitives.py:113-127: params = (center, radius)."""
        return self._run(_SPHERE, points, _flat(params), sqrt)

    def distance_from_cylinder(self, points, params, sqrt=False):
        """src/primitives.py:129-161: params = (axis, center, radius)."""
        return self._run(_CYLINDER, points, _flat(params), sqrt)

    def distance_from_cone(self, points, params, sqrt=False):
        """src/primitives.py:166-195: params = (apex, axis, theta)."""
        return self._run(_CONE, points, _flat(params), sqrt)


class ResidualLoss:
    """src/primitives.py:18-44 (analytic primitives; spline distances are outside the path)."""

    def __init__(self, reduce=True, one_side=False):
        cp = ComputePrimitiveDistance(reduce, one_side=one_side)
        self.routines = {"torus": cp.distance_from_torus, "sphere": cp.distance_from_sphere,
                         "cylinder": cp.distance_from_cylinder, "cone": cp.distance_from_cone,
                         "plane": cp.distance_from_plane}

    def residual_loss(self, Points, parameters, sqrt=False):
        distances = {}
        for k, v in parameters.items():
            if v is None:
                continue
            distances[k] = [v[0], self.routines[v[0]](points=Points[k], params=v[1:], sqrt=sqrt)]
        return distances
