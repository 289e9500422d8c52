def collect_exp_stats(*a, **k):
    return None
