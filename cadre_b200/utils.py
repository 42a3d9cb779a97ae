"""In-process stand-ins for the reference's cross-process sync primitives (ppo_agent/utils.py:31-70, 108-126) and
the `chief` entry point (ppo_agent/chief.py:8-27). With one process per GPU and an NCCL all-reduce the 1 Hz polling
parameter server disappears; these keep the names so that code written against the reference still imports."""


class Counter:
    """ppo_agent/utils.py:31-70 (mp.Value + mp.Lock there; a plain int here: single process per GPU)."""

    def __init__(self, val=True):
        self.val = 0

    def get(self):
        return self.val

    def increment(self):
        self.val += 1

    def reset(self):
        self.val = 0


class TrafficLight:
    """ppo_agent/utils.py:108-126."""

    def __init__(self, val=True):
        self.val = False

    def get(self):
        return self.val

    def reset(self):
        self.val = False

    def switch(self):
        self.val = not self.val


def chief(update_threshold, traffic_light, counter, shared_model_list, shared_grad_buffers, optimizer,
          son_process_counter=None, max_grad_norm=250.0, total_thread=1):
    """One pass of the chief body (chief.py:12-24) on the flat buffers: when `counter` has reached
    `update_threshold`, apply per-module clip + Adam (the CUDA kernels of `optimizer`, a
    cadre_b200.learner.Learner) to the SUM held in `shared_grad_buffers`, reset, flip the light.
    Returns True if a step was applied. (The reference loops forever with a 1 s sleep; callers loop.)"""
    if counter.get() < update_threshold:
        return False
    optimizer.grads.copy_(shared_grad_buffers.grads)
    optimizer.step_count += 1
    optimizer.engine.adam_step(optimizer.params, optimizer.grads, optimizer.exp_avg, optimizer.exp_avg_sq,
                               optimizer.step_count, max_grad_norm, optimizer.lr)
    shared_grad_buffers.reset()
    counter.reset()
    traffic_light.switch()
    return True
