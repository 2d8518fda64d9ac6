"""Importable but unused on the hot path."""


class WeightDecay:
    def __init__(self, rate):
        self.rate = rate
