"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the resample front-end, SURVEY 8f row 3.

The reference resamples inside ``librosa.load(path, sr=8000)`` (create_train_dataset.py:204,225; create_test_dataset.py:139;
test.py:80): ``librosa.to_mono`` (mean over channels) followed by ``librosa.resample`` with the 0.10 default ``res_type='soxr_hq'``.
soxr is a third-party dependency (requirements.txt: librosa==0.10.2.post1 -> soxr) that is absent from /root/reference and from
this image, and the reference fixes no filter of its own, so PARITY WITH THE REFERENCE'S RESAMPLER IS UNPINNED.  The oracle is the
published polyphase algorithm in scipy.signal.resample_poly (float64, default Kaiser beta = 5 window), which produces the same
number of output samples, ceil(L * target / orig), and like soxr is a linear-phase low-pass at the new Nyquist."""
import math

import numpy as np
from scipy import signal


def to_mono(y: np.ndarray) -> np.ndarray:
    """librosa.to_mono: mean over the leading (channel) axis of a channel-first array."""
    y = np.asarray(y)
    return y if y.ndim == 1 else np.mean(y, axis=tuple(range(y.ndim - 1)))


def resample(y: np.ndarray, orig_sr: int, target_sr: int) -> np.ndarray:
    """float64 polyphase resampling of the last axis."""
    g = math.gcd(int(orig_sr), int(target_sr))
    up, down = int(target_sr) // g, int(orig_sr) // g
    y = np.asarray(y, dtype=np.float64)
    if up == down:
        return y.copy()
    return signal.resample_poly(y, up, down, axis=-1)


def load_decoded(audio: np.ndarray, native_sr: int, sr: int = 8000) -> np.ndarray:
    """to_mono + resample, the numeric tail of librosa.load(path, sr=sr)."""
    return resample(to_mono(np.asarray(audio, dtype=np.float32)).astype(np.float64), native_sr, sr)
