"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's slide preprocessing.

Follows /root/reference/celldetection_scripts/cpn_inference.py:196-222 (``preprocess``) and
/root/reference/celldetection/data/misc.py:156-161 (``normalize_percentile``) with whole-image numpy passes, exactly as the
reference runs them on the CPU.  Third-party calls restated from their published sources (packages absent here -> parity
of these three lines is unpinned): ``skimage.img_as_ubyte`` (float [0, 1] -> ``rint(x * 255)`` in float64),
``albumentations.augmentations.functional.gamma_transform`` / ``brightness_contrast_adjust`` for uint8 images (both are
``cv2.LUT`` look-ups; ``cv2.LUT`` itself is plain indexing)."""
import numpy as np


def normalize_percentile(image, percentile=99.9, to_uint8=True):
    """data/misc.py:156-161"""
    if not isinstance(percentile, (list, tuple)):
        percentile = (100 - percentile, percentile)
    low, high = np.percentile(image, percentile)
    img = (np.clip(image, low, high) - low) / (high - low)
    return np.clip(np.rint(img * 255.), 0, 255).astype(np.uint8) if to_uint8 else img


def gamma_transform(img, gamma):
    table = (np.arange(0, 256.0 / 255, 1.0 / 255) ** gamma) * 255
    return table.astype(np.uint8)[img]


def brightness_contrast_adjust(img, alpha=1., beta=0.):
    lut = np.arange(0, 256).astype(np.float32)
    if alpha != 1:
        lut *= alpha
    if beta != 0:
        lut += beta * np.mean(img)
    return np.clip(lut, 0, 255).astype(np.uint8)[img]


def rgb2gray(img):
    """cv2.cvtColor(img, COLOR_RGB2GRAY | COLOR_RGBA2GRAY) for uint8 (OpenCV 4.x 15-bit fixed point; tests pin it against the
    installed cv2)."""
    r, g, b = (img[..., i].astype(np.int64) for i in range(3))
    return ((9798 * r + 19235 * g + 3735 * b + (1 << 14)) >> 15).astype(np.uint8)


def preprocess(img, gamma=1., contrast=1., brightness=0., percentile=None, grayscale=False):
    """cpn_inference.py:196-222 (grayscale: the 3- and 4-channel branches)."""
    if percentile is not None:
        img = normalize_percentile(img, percentile)
    if img.itemsize > 1:
        img = normalize_percentile(img)
    if grayscale and img.ndim == 3:
        img = img.squeeze(-1) if img.shape[-1] == 1 else rgb2gray(img)
    if img.ndim == 2:
        img = np.repeat(img[..., None], 3, axis=-1)        # cv2.COLOR_GRAY2RGB
    if gamma != 1.:
        img = gamma_transform(img, gamma)
    if contrast != 1.:
        img = brightness_contrast_adjust(img, alpha=contrast, beta=brightness)
    return img
