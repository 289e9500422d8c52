class Planetoid:  # name only
    def __init__(self, *a, **k):
        raise RuntimeError("torch_geometric shim: Planetoid loading is not available offline")
