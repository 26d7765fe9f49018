"""skimage.filters.gaussian for 2-D float images = scipy.ndimage.gaussian_filter (skimage's
implementation calls exactly this after converting to float; preserve_range only disables the
[0,1] rescale of integer inputs)."""
import numpy as np
import scipy.ndimage as ndi


def gaussian(image, sigma=1, output=None, mode='nearest', cval=0, preserve_range=False, truncate=4.0,
             *, channel_axis=None, **_ignored):
    img = np.asarray(image, dtype=np.float64)
    return ndi.gaussian_filter(img, sigma, output=output, mode=mode, cval=cval, truncate=truncate)
