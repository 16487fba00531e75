"""gym.spaces stand-in (Space, Discrete, Box, Dict) - test infrastructure only."""
import random as _random


class Space:
    def sample(self):
        raise NotImplementedError


class Discrete(Space):
    def __init__(self, n):
        self.n = int(n)

    def sample(self):
        return _random.randrange(self.n)

    def contains(self, x):
        return 0 <= int(x) < self.n


class Box(Space):
    def __init__(self, low, high, shape=None, dtype='float32'):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape or ()), dtype


class Dict(Space):
    def __init__(self, spaces=None):
        self.spaces = dict(spaces or {})

    def __getitem__(self, key):
        return self.spaces[key]

    def sample(self):
        return {k: s.sample() for k, s in self.spaces.items()}
