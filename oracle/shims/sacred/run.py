class Run:  # name only (type annotation in src/trainers/bilevel.py:6)
    def log_scalar(self, *a, **k):
        pass
