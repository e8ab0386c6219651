


def invalidate_kernel_caches(module):
    """Forget every folded / packed weight copy and every captured CUDA graph that the kernel launchers under ``module``
    hold.  The launchers key their prepared workspaces on (data_ptr, _version) of the parameters, which in-place updates
    through ``p.data`` (EMA, custom weight loaders) do not change: call this after such an update; the next call of each
    operator runs its ``sfb_*_prepare`` again.  (``load_state_dict``, optimizer steps and ``.to()`` are seen without it.)"""
    n = 0
    for m in module.modules():
        for name in ('_engine', '_enc_tail_engine', '_transition_engine'):
            eng = m.__dict__.get(name)
            if eng is not None and hasattr(eng, 'invalidate'):
                eng.invalidate()
                n += 1
        if m.__dict__.pop('_loop_graphs', None) is not None:
            n += 1
    return n

