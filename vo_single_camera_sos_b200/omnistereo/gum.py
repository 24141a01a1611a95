"""Mirror of the hot-path part of omnistereo/gum.py: the Generalised Unified Model of one mirror and the GUM stereo rig.
The calibration / Jacobian half of the reference file (gum.py:474-2500) is out of scope."""
import numpy as np

from . import device_context, to_device
from .camera_models import OmniCamModel, OmniStereoModel


class Parameters(object):
    """Pre-calibration parameters (gum.py:48-317), same attribute names as the reference."""

    def __init__(self, precalib_filename=None, new_method=True, cam_model=None, **kwargs):
        self.new_method = new_method
        self.precalib_filename = precalib_filename
        z_axis = cam_model.z_axis if cam_model is not None else 1.0
        self.xi1, self.xi2, self.xi3 = 0.0, 0.0, 1.0 * z_axis
        self.k1 = self.k2 = self.k3 = self.p1 = self.p2 = 0.0
        self.use_distortion = True
        self.l1 = self.l2 = self.l3 = 0.0
        if "image_size_pixels" in kwargs:
            self.image_size = np.array(kwargs.get("image_size_pixels"))
        if "center_uv_point" in kwargs:
            self.center_point = np.array(kwargs.get("center_uv_point"), float)
        else:
            self.center_point = (self.image_size / 2.0) - 1
        self.u_center, self.v_center = self.center_point
        self.center_point_inner = self.center_point_outer = None
        self.gamma1 = self.gamma2 = 300.0
        self.alpha_c = 0.0
        self.roi_min_x = self.roi_min_y = self.roi_max_x = self.roi_max_y = None
        self.set_gum_params()
        self.set_generalized_cam_params()

    def set_gum_params(self, **kwargs):
        """gum.py:169-200."""
        if "center_uv_point" in kwargs:
            self.center_point = np.array(kwargs.get("center_uv_point"), float)
            self.u_center, self.v_center = self.center_point
        for k in ("xi1", "xi2", "xi3"):
            if k in kwargs:
                setattr(self, k, kwargs[k])
        self.Cp = np.array([self.xi1, self.xi2, self.xi3], float)
        self.set_generalized_cam_params(**{k: v for k, v in kwargs.items() if k in ("alpha_c", "gamma1", "gamma2", "u_center", "v_center")})

    def set_generalized_cam_params(self, **kwargs):
        """gum.py:121-140: generalised camera matrix and its inverse."""
        for k in ("alpha_c", "gamma1", "gamma2", "u_center", "v_center"):
            if k in kwargs:
                setattr(self, k, kwargs[k])
        self.center_point = np.array([self.u_center, self.v_center], float)
        self.inv_K11 = 1 / self.gamma1
        self.inv_K12 = -self.alpha_c / self.gamma2
        self.inv_K13 = self.alpha_c * self.v_center / self.gamma2 - self.u_center / self.gamma1
        self.inv_K22 = 1 / self.gamma2
        self.inv_K23 = -self.v_center / self.gamma2


class GUM(OmniCamModel):
    """Single-mirror Generalised Unified Model (gum.py:320-383, 2512-2940)."""

    def __init__(self, precalib_filename=None, new_method=True, z_axis=1.0, use_theoretical_xi_and_gamma=False, **kwargs):
        self.new_method = new_method
        self.z_axis = z_axis
        self.mirror_name, self.mirror_number = ("top", 1) if z_axis > 0 else ("bottom", 2)
        self.precalib_filename = precalib_filename
        self._init_default_values()
        self.precalib_params = Parameters(precalib_filename, new_method=new_method, cam_model=self, **kwargs)
        self.set_model_params()

    def set_model_params(self, **kwargs):
        """gum.py:361-383: projection centre Cp and the z of the normalised projection plane."""
        self.precalib_params.set_gum_params(**kwargs)
        self.Pm = np.array([0., 0., 0.])
        self.Cp_wrt_M = self.precalib_params.Cp
        self.plane_n = np.array([0, 0, 1.0 * self.z_axis])
        self.plane_k = self.Cp_wrt_M[2] - self.plane_n[2]

    def gum_vector(self):
        p = self.precalib_params
        return np.array([p.xi1, p.xi2, p.xi3, p.k1, p.k2, p.k3, p.gamma1, p.gamma2, p.alpha_c, p.u_center, p.v_center,
                         p.l1, p.l2, p.l3, p.p1, p.p2, self.Cp_wrt_M[2] - self.z_axis, float(bool(p.use_distortion))],
                        np.float64)

    def get_center(self):
        return self.precalib_params.u_center, self.precalib_params.v_center

    # ---- F3 -----------------------------------------------------------------------------------------------------------
    def get_pixel_from_3D_point_wrt_M(self, Pw_wrt_M, visualize=False):
        """(u, v, m_homo) of points given wrt the GUM frame [M] (gum.py:2512-2551), shapes as the reference."""
        P = np.asarray(Pw_wrt_M, np.float64)
        shape = P.shape[:-1]
        uv = device_context().gum_project(self.gum_vector(), to_device(np.ascontiguousarray(P[..., :3].reshape(-1, 3))))
        uv = uv.cpu().numpy().reshape(shape + (2,))
        u, v = uv[..., 0], uv[..., 1]
        return u, v, np.dstack((u, v, np.ones_like(u)))

    def get_3D_point_from_angles_wrt_focus(self, azimuth, elevation):
        """gum.py:2564-2575."""
        return self.map_angles_to_unit_sphere(elevation, azimuth)

    # ---- F9 -----------------------------------------------------------------------------------------------------------
    def lift_pixel_to_unit_sphere_wrt_focus(self, m, visualize=False, debug=False, return_angles=False):
        """Point(s) on the unit sphere for omni-image pixel(s) m[..., :2] (gum.py:2673-2940, new_method branch)."""
        if not self.new_method:
            raise NotImplementedError("only the new (Zhang) GUM lifting is mirrored")
        m = np.asarray(m, np.float64)
        shape = m.shape[:-1]
        sphere, az, el = device_context().lift_gum(self.gum_vector(), to_device(np.ascontiguousarray(m[..., :2].reshape(-1, 2))))
        Ps = sphere.cpu().numpy().reshape(shape + (3,))
        if return_angles:
            return Ps, az.cpu().numpy().reshape(shape), el.cpu().numpy().reshape(shape)
        return Ps

    def set_elevation_limits_from_radii(self):
        """camera_models.py:1312-1382, 1402-1456: lift 360 pixels on each radial bound; the top mirror sees its highest
        elevation at the outer radius, the bottom mirror at the inner one."""
        p = self.precalib_params
        phi = np.linspace(0, 2 * np.pi, num=360, endpoint=False)
        r_low, r_high = ((self.inner_img_radius, self.outer_img_radius) if self.mirror_number == 1
                         else (self.outer_img_radius, self.inner_img_radius))
        c_low = p.center_point_inner if self.mirror_number == 1 else p.center_point_outer
        c_high = p.center_point_outer if self.mirror_number == 1 else p.center_point_inner
        c_low = p.center_point if c_low is None else c_low
        c_high = p.center_point if c_high is None else c_high
        low = np.stack([c_low[0] + r_low * np.cos(phi), c_low[1] + r_low * np.sin(phi)], 1)
        high = np.stack([c_high[0] + r_high * np.cos(phi), c_high[1] + r_high * np.sin(phi)], 1)
        self.lowest_elevation_angle = float(np.nanmin(self._elevations_by_inverting_the_forward_projection(low)))
        self.highest_elevation_angle = float(np.nanmax(self._elevations_by_inverting_the_forward_projection(high)))

    def _elevations_by_inverting_the_forward_projection(self, m):
        """One-off set-up helper.  The reference finds these boundary angles by optimising the FORWARD projection per
        pixel (get_direction_angles_from_pixel_using_forward_projection, camera_models.py:1343,1372) because the closed
        form inverse distortion is only approximate; the forward radial model rho_d = rho_u (1 + k1 rho_u^2 + k2 rho_u^4
        + k3 rho_u^6) is inverted here exactly by Newton's method and the undistorted point is lifted by the device
        kernel with distortion switched off."""
        p = self.precalib_params
        xd = p.inv_K11 * m[:, 0] + p.inv_K12 * m[:, 1] + p.inv_K13
        yd = p.inv_K22 * m[:, 1] + p.inv_K23
        rd = np.hypot(xd, yd)
        ru = rd.copy()
        if p.use_distortion:
            for _ in range(30):
                r2 = ru * ru
                f = ru * (1 + p.k1 * r2 + p.k2 * r2 ** 2 + p.k3 * r2 ** 3) - rd
                df = 1 + 3 * p.k1 * r2 + 5 * p.k2 * r2 ** 2 + 7 * p.k3 * r2 ** 3
                ru = ru - f / df
        s = np.where(rd > 0, ru / np.where(rd > 0, rd, 1.0), 1.0)
        xu, yu = xd * s, yd * s
        u = p.gamma1 * xu + p.gamma1 * p.alpha_c * yu + p.u_center
        v = p.gamma2 * yu + p.v_center
        g = self.gum_vector()
        g[-1] = 0.0  # use_distortion off: (u, v) already encode the undistorted point
        _, _, el = device_context().lift_gum(g, to_device(np.ascontiguousarray(np.stack([u, v], 1))))
        return el.cpu().numpy()


class GUMStereo(OmniStereoModel):
    """The vertically folded omnistereo rig of two GUMs (gum.py:3043-3116)."""

    def get_baseline(self):
        return self.top_model.F[2, 0] - self.bot_model.F[2, 0]
