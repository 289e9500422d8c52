class TelegramObserver:  # name only (src/utils/tracking.py:6)
    def __init__(self, *a, **k):
        pass


class MongoObserver:
    def __init__(self, *a, **k):
        pass
