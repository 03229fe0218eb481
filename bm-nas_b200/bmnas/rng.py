"""Seeds for the in-kernel Philox dropout streams (one independent stream per launch plan)."""
_state = {'seed': 0x1234ABCD5678EF01, 'n': 0}


def manual_seed(seed):
    _state['seed'] = int(seed) & 0x7FFFFFFFFFFFFFFF
    _state['n'] = 0


def next_seed():
    _state['n'] += 1
    return (_state['seed'] + _state['n'] * 0x9E3779B97F4A7C15) & 0x7FFFFFFFFFFFFFFF
