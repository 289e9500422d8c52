class GAE:  # name only (src/trainers/pretrainer.py:8)
    def __init__(self, *a, **k):
        raise RuntimeError("torch_geometric shim: GAE is not available")
