"""Mirror of omnistereo/panorama.py for the hot path: LUT generation, LUT remap, panorama pixel <-> direction angles."""
import cv2
import numpy as np
import torch

from . import device_context, to_device


class Panorama(object):
    """Cylindrical panorama of one mirror view.  Attribute names follow the reference (panorama.py:51-87) so that
    pickled instances keep working: world2cam_LUT_map_x/y (float64 rows x cols, NaN outside the mirror's own FOV),
    rows, cols, pixel_size, cyl_radius, cyl_height_max, z_height_min, cyl_circumference, panoramic_img, ..."""

    def __init__(self, projection_model, **kwargs):
        self.model = projection_model
        self.name = getattr(projection_model, "mirror_name", "") + " " + __name__
        self.omni_img = None
        self.panoramic_img = None
        self.psi_LUT = None
        self.theta_LUT = None
        self.theta_LUT_validated = None
        self.world2cam_LUT_map_x = None
        self.world2cam_LUT_map_y = None
        self.cyl_radius = 1.0
        self.globally_highest_elevation_angle = self.model.globally_highest_elevation_angle
        self.globally_lowest_elevation_angle = self.model.globally_lowest_elevation_angle
        self.pixel_size = 1.0
        self._set_cylinder_height()
        self.azimuthal_masks = []
        self.set_panorama_dimensions(**kwargs)

    # ---- geometry (panorama.py:89-172) ------------------------------------------------------------------------
    def set_panorama_dimensions(self, **kwargs):
        self.interpolation_method = kwargs.get("interpolation", cv2.INTER_LINEAR)
        self.border_method = kwargs.get("border_method", cv2.BORDER_CONSTANT)
        self.azimuthal_shift = kwargs.get("azimuthal_shift", 0)
        self.width = kwargs.get("width", 600)
        self.cols = int(np.ceil(self.width))
        self._resolve_dimensions_pixel_sizing()
        self._generate_LUTs()

    def _set_cylinder_height(self):
        self.cyl_height_max = self.cyl_radius * np.tan(self.globally_highest_elevation_angle)
        self.z_height_min = self.cyl_radius * np.tan(self.globally_lowest_elevation_angle)
        self.cyl_height = self.cyl_height_max - self.z_height_min

    def _resolve_dimensions_pixel_sizing(self):
        self.cyl_circumference = 2 * np.pi * self.cyl_radius
        self.pixel_size = self.cyl_circumference / float(self.cols)
        self.height = self.cyl_height / self.pixel_size
        self.rows = int(np.ceil(self.height))
        self.aspect_ratio = float(self.cols) / float(self.rows)

    def regenerate_LUTs(self, method="pixel sizing"):
        self._resolve_dimensions_pixel_sizing()
        self._generate_LUTs()

    # ---- F3: LUT generation (panorama.py:414-492) on the device -------------------------------------------------
    def _generate_LUTs(self):
        ctx = device_context()
        self.psi_LUT = np.linspace(0, 2 * np.pi, num=self.cols, endpoint=False)[::-1].copy()
        cyl_height_LUT = np.linspace(self.cyl_height_max, self.z_height_min, num=self.rows, endpoint=False)
        self.theta_LUT = np.arctan2(cyl_height_LUT, self.cyl_radius)
        lo, hi = self.model.lowest_elevation_angle, self.model.highest_elevation_angle
        self.theta_LUT_validated = np.where((lo <= self.theta_LUT) & (self.theta_LUT <= hi), self.theta_LUT, np.nan)
        mx, my = ctx.lut_build(self.model.gum_vector(), self.rows, self.cols, self.cyl_height_max, self.z_height_min, lo, hi)
        self.world2cam_LUT_map_x = mx.cpu().numpy()
        self.world2cam_LUT_map_y = my.cpu().numpy()
        self._drop_device_state()

    def _drop_device_state(self):
        self.__dict__.pop("_dev", None)

    def __getstate__(self):  # device handles are rebuilt lazily after unpickling
        d = dict(self.__dict__)
        d.pop("_dev", None)
        return d

    def _device_lut(self, src_hw, mask):
        """Packed fixed-point LUT on the device, cached per (source shape, mask identity, LUT identity)."""
        key = (tuple(src_hw), id(mask), id(self.world2cam_LUT_map_x), id(self.world2cam_LUT_map_y))
        dev = self.__dict__.get("_dev")
        if dev is None or dev[0] != key:
            ctx = device_context()
            lut = ctx.lut_pack(to_device(self.world2cam_LUT_map_x), to_device(self.world2cam_LUT_map_y), src_hw,
                               mask=None if mask is None else to_device(mask))
            dev = (key, lut, mask)
            self.__dict__["_dev"] = dev
        return dev[1]

    # ---- F1 (+F2): remap (panorama.py:258-321) -------------------------------------------------------------------
    def get_panoramic_image(self, input_omni_img, set_own=True, crop_out_bottom=False, border_RGB_color=None,
                            use_floating_point_prec=True, mask=None, mask_BGR_color=(0, 0, 0)):
        """Panoramic representation of `input_omni_img` — bit-exact with the reference's cv2.remap(INTER_LINEAR,
        BORDER_CONSTANT) call.  `mask` / `mask_BGR_color` (an extension) fold the mirror mask of
        get_fully_masked_images into the same kernel: masked-out source pixels read as `mask_BGR_color`."""
        border = (0, 0, 0) if border_RGB_color is None else (border_RGB_color[2], border_RGB_color[1], border_RGB_color[0])
        img = np.ascontiguousarray(input_omni_img)
        if img.dtype != np.uint8:
            raise TypeError("the panorama remap kernel handles 8-bit images")
        if set_own:
            self.omni_img = input_omni_img
        self.color_channels = img.ndim
        ctx = device_context()
        lut = self._device_lut(img.shape[:2], mask)
        ch = 1 if img.ndim == 2 else img.shape[2]
        out = ctx.remap(to_device(img[None]), lut, border=tuple(border)[:ch] if ch > 1 else (border[0],),
                        background=tuple(mask_BGR_color)[:ch] if ch > 1 else (mask_BGR_color[0],))
        panoramic_img = out[0, 0].cpu().numpy()
        if set_own:
            self.panoramic_img = panoramic_img
        return panoramic_img

    def set_panoramic_image(self, omni_img, idx=-1, view=True, win_name_modifier="", border_RGB_color=None, **kwargs):
        return self.get_panoramic_image(omni_img, set_own=True, border_RGB_color=border_RGB_color, **kwargs)

    # ---- F7: panorama pixel -> direction angles (panorama.py:616-666) ---------------------------------------------
    def pano_vector(self):
        return np.array([self.cols, self.rows, self.pixel_size, self.cyl_height_max, self.cyl_circumference,
                         self.cyl_radius], np.float64)

    def get_direction_angles_from_pixel_pano(self, m_pano, use_LUTs=False):
        """azimuth, elevation (float64, NaN outside the panorama) of panorama pixels m_pano[..., :2]."""
        m = np.asarray(m_pano, np.float64)
        shape = m.shape[:-1]
        flat = np.ascontiguousarray(m[..., :2].reshape(-1, 2))
        if flat.shape[0] == 0:
            return np.zeros(shape), np.zeros(shape)
        az, el, _ = device_context().lift_pano_f64(self.pano_vector(), to_device(flat))
        return az.cpu().numpy().reshape(shape), el.cpu().numpy().reshape(shape)

    def get_azimuth_from_panorama_col_without_LUT(self, col):
        col = np.asarray(col, np.float64)
        m = np.stack([col, np.zeros_like(col)], axis=-1)
        return self.get_direction_angles_from_pixel_pano(m)[0]

    def get_elevation_from_panorama_row_without_LUT(self, row):
        row = np.asarray(row, np.float64)
        m = np.stack([np.zeros_like(row), row], axis=-1)
        return self.get_direction_angles_from_pixel_pano(m)[1]

    def get_panorama_col_from_azimuth(self, azimuth):
        return int(np.floor((self.cyl_circumference - np.mod(azimuth, 2 * np.pi)) / self.pixel_size)) % self.cols

    def generate_azimuthal_masks(self, azimuth_mask_degrees, overlap_degrees=0, mask_also_on_elev=True, elev_mask_padding=0,
                                 stand_masks_azimuth_coord_in_degrees_list=(), stand_masks_width_in_degrees=1, show=False):
        """Bucket masks for feature detection (panorama.py:520-589); host-side setup, rectangles only."""
        self.azimuthal_masks = []
        step = np.deg2rad(azimuth_mask_degrees)
        valid_rows = ~np.isnan(self.theta_LUT_validated)
        stands = np.ones(self.cols, bool)
        half = np.deg2rad(stand_masks_width_in_degrees / 2.0)
        for a in np.deg2rad(np.asarray(list(stand_masks_azimuth_coord_in_degrees_list), float)):
            az = self.psi_LUT
            stands &= ~((az >= a - half) & (az <= a + half))
        for d in np.arange(0, 2 * np.pi - step / 2.0, step):
            cols = (self.psi_LUT >= d) & (self.psi_LUT < d + step) & stands
            mask = np.zeros((self.rows, self.cols), np.uint8)
            mask[np.ix_(valid_rows if mask_also_on_elev else np.ones(self.rows, bool), cols)] = 255
            self.azimuthal_masks.append(mask)
