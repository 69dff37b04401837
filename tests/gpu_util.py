def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def rel_l2(got, want):
    import numpy as np
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    d = np.linalg.norm(want.ravel())
    return float(np.linalg.norm((got - want).ravel()) / d) if d else float(np.linalg.norm(got.ravel()))
