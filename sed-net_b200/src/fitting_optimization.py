"""FittingModule with the reference's method names and parameter layouts (reference src/fitting_optimization.py:118-245)
for the analytic primitives; every fit runs on the batched fit kernel of libsednet_b200.so (``src.primitive_forward.Fit``).

``fitting.parameters[ids]`` is filled exactly as the reference leaves it (shapes in brackets):
    ["plane", axis (3,1), distance ()]                    :160-167
    ["cone", apex (1,3), axis (3,1), theta ()]            :184-196
    ["cylinder", a (3,1), center (1,3), radius ()]        :208-215
    ["sphere", center (1,3), radius ()]                   :230-237
The SplineNet passes (:135-158) and the ``sample_points=True`` mesh sampling belong to the reference's spline / meshing
code, outside this path: they raise."""
from .primitive_forward import Fit


class FittingModule:
    def __init__(self, closed_splinenet_path=None, open_splinenet_path=None):
        self.fitting = Fit()
        self.closed_splinenet_path = closed_splinenet_path
        self.open_splinenet_path = open_splinenet_path

    @staticmethod
    def _no_sampling(sample_points):
        if sample_points:
            raise NotImplementedError("sample_points=True samples a mesh of the fitted primitive (visualisation path of the "
                                      "reference, src/fitting_optimization.py:168-182); only the fits are on the hot path")

    def forward_pass_open_spline(self, points, ids, weights, if_optimize=False):
        raise NotImplementedError("SplineNet patches are outside the analytic-primitive hot path")

    def forward_pass_closed_spline(self, points, ids, weights, if_optimize=False):
        raise NotImplementedError("SplineNet patches are outside the analytic-primitive hot path")

    def forward_pass_plane(self, points, normals, weights, ids, sample_points=False):
        self._no_sampling(sample_points)
        axis, distance = self.fitting.fit_plane_torch(points=points, normals=normals, weights=weights, ids=ids)
        self.fitting.parameters[ids] = ["plane", axis.reshape((3, 1)), distance]
        return None

    def forward_pass_cone(self, points, normals, weights, ids, sample_points=False):
        self._no_sampling(sample_points)
        apex, axis, theta = self.fitting.fit_cone_torch(points, normals, weights=weights, ids=ids)
        self.fitting.parameters[ids] = ["cone", apex.reshape((1, 3)), axis.reshape((3, 1)), theta]
        return None

    def forward_pass_cylinder(self, points, normals, weights, ids, sample_points=False):
        self._no_sampling(sample_points)
        a, center, radius = self.fitting.fit_cylinder_torch(points, normals, weights, ids=ids)
        self.fitting.parameters[ids] = ["cylinder", a, center, radius]
        return None

    def forward_pass_sphere(self, points, normals, weights, ids, sample_points=False):
        self._no_sampling(sample_points)
        center, radius = self.fitting.fit_sphere_torch(points, normals, weights, ids=ids)
        self.fitting.parameters[ids] = ["sphere", center, radius]
        return None
