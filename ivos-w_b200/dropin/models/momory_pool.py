"""Minimal stand-in for the reference's models/momory_pool.py::ReplayMemory (ring buffer only).
The pandas/CSV mirror (load_from_csv / push_to_csv, momory_pool.py:44-153) is training-time
bookkeeping outside the scoring path (OUT OF SCOPE) and is not reproduced."""
import random


class ReplayMemory(object):
    def __init__(self, capacity):
        self.capacity = capacity
        self.memory = []
        self.position = 0

    def push(self, *args):
        if len(self.memory) < self.capacity:
            self.memory.append(None)
        self.memory[self.position] = args
        self.position = (self.position + 1) % self.capacity

    def random_sample(self, batch_size):
        return random.sample(self.memory, min(batch_size, len(self.memory)))

    def __len__(self):
        return len(self.memory)
