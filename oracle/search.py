"""CPU restatement of FREUD's per-feature max-activation search.

TEST INFRASTRUCTURE ONLY.  numpy; follows src/utils/activations.py:19-132 and
src/dataset/activations.py:116-174.  ``n_frames`` replaces the audio decode of
trim_activation (utils/activations.py:19-29) by its own arithmetic on a
num_samples table: int((num_samples / sample_rate) / TIMESTEP_S), python floats.
"""
import numpy as np

TIMESTEP_S = 30 / 1500  # src/utils/constants.py:17
SAMPLE_RATE = 16000  # src/utils/constants.py:6


def n_frames_from_samples(num_samples, sample_rate=SAMPLE_RATE):
    """utils/activations.py:26-28."""
    return int((num_samples / sample_rate) / TIMESTEP_S)


def dense_from_indexed(vals, idx, feature_idx):
    """utils/activations.py:41-57: out[i,j] = vals[i,j,pos] where idx[i,j,pos]==feature
    (first match; at most one per frame for top-k output), else 0.  fp32 output."""
    hit = idx == feature_idx
    any_hit = hit.any(-1)
    pos = hit.argmax(-1)
    g = np.take_along_axis(vals, pos[..., None], -1)[..., 0]
    return np.where(any_hit, g, 0).astype(np.float32)


def top_activations(acts, filenames, n_frames, n_files, max_val, min_val, absolute_magnitude,
                    return_max_per_file):
    """utils/activations.py:61-132 on one feature's traces.

    acts: [N_files, T] array (the feature column, or dense_from_indexed output).
    Returns (pq, max_per_file): pq[i] = (filename, trimmed trace, value, time).
    Ranking = stable descending sort by value after every append, truncated to
    n_files (ties keep the earlier file), exactly as :129-130.
    """
    pq, max_per_file = [], []

    def allowed(v):
        if max_val is not None and v > max_val:
            return False
        if min_val is not None and v < min_val:
            return False
        return True

    for i, fname in enumerate(filenames):
        a = acts[i][: n_frames[i]]
        if absolute_magnitude:
            j = int(np.argmax(np.abs(a)))
            signed = float(a[j])
            ok = allowed(signed)
            value = abs(signed)
            if return_max_per_file:
                max_per_file.append(signed)
        else:
            value = float(a.max())
            ok = allowed(value)
            if return_max_per_file:
                max_per_file.append(value)
        if ok:
            loc = int(np.argmax(a))  # signed argmax even in abs mode (:119)
            pq.append((fname, a, value, loc * TIMESTEP_S))
            pq.sort(key=lambda t: t[2], reverse=True)
            pq = pq[:n_files]
    return pq, (max_per_file if return_max_per_file else None)
