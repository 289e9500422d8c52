def create_mongodb_observer(*a, **k):
    return None
