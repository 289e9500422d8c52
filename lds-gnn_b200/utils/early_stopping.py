"""Early stopping with the reference's rule (src/utils/early_stopping.py:7-39): keep going while
`step <= patience` or the new value is <= the mean of the previous `patience` values; remembers the
parameters handed in with the last accepted update. Host logic only."""
import numpy as np


class EarlyStopping:
    def __init__(self, patience: int, max_epochs: int = 10000):
        self.patience = patience
        self.max_epochs = max_epochs
        self.abort = False
        self.curr_step = 0
        self.losses = []
        self.model_state_dict = None
        self.model_params = None

    def update(self, new_value, model=None, model_params=None):
        self.losses.append(new_value)
        window = self.losses[-(self.patience + 1):-1]
        if self.curr_step <= self.patience or new_value <= np.mean(window):
            if model is not None:
                self.model_state_dict = model.state_dict()
            if model_params is not None:
                self.model_params = model_params
        else:
            self.abort = True
        if self.curr_step >= self.max_epochs:
            self.abort = True
        self.curr_step += 1

    def best_model_state_dict(self):
        return self.model_state_dict
