""" Context managers (reference romcomma/user/contexts.py:32-82): ``Timer`` and ``Environment``.
``Environment`` selects the CUDA device instead of a TensorFlow logical device; float64 is the only float."""
from __future__ import annotations

from contextlib import contextmanager
from datetime import timedelta
from time import time

import torch

from romcomma.base.definitions import *


@contextmanager
def Timer(name: str = '', is_inline: bool = True):
    """ Times the enclosed block and prints ``Running <name> took h:mm:ss``. An empty name is silent."""
    start = time()
    if name != '':
        print(f'Running {name}', end='' if is_inline else '...\n', flush=True)
    yield
    if name != '':
        print(f'{" " if is_inline else "..."}took {timedelta(seconds=int(time() - start))}.')


@contextmanager
def Environment(name: str = '', device: str = '', **kwargs):
    """ Sets up the environment to run operations.

    Args:
        name: Printed as what is being run.
        device: ``'GPU'`` / ``'GPU:3'`` / ``'cuda:3'`` select a CUDA device; anything else keeps the current one. There is no CPU device on this path.
        **kwargs: ``float`` must be float64 if given; ``eager`` is accepted and ignored (there is no graph compiler here).
    """
    with Timer(name):
        kwargs = kwargs | {'float': 'float64'}
        kwargs.pop('eager', None)
        print(' using B200 kernels(' + ', '.join(f'{k}={v!r}' for k, v in kwargs.items()), end=')')
        index = None
        tail = device[max(device.rfind('GPU'), device.rfind('cuda')):] if ('GPU' in device or 'cuda' in device) else ''
        if ':' in tail:
            index = int(tail.split(':')[1])
        elif tail:
            index = torch.cuda.current_device() if torch.cuda.is_available() else 0
        print(f' on cuda:{index}...' if index is not None else '...')
        if index is not None and torch.cuda.is_available():
            with torch.cuda.device(index):
                yield
        else:
            yield
        print('...Running ' + name, end='')
