"""Task-level parallelism for independent fits (role of GPErks/utils/concurrency.py): a process pool with the
'spawn' start method (CUDA contexts do not survive fork), or in-process execution when one worker suffices."""
import concurrent.futures
import multiprocessing

_WORKER_DEVICE = None


def _bind_worker(device_queue, n_workers=1):
    """Pool initializer: every worker process takes ONE device for its whole life, so at most ``len(devices)`` fits run at a
    time and never two on the same GPU (job-indexed devices let a free worker start a second fit on a busy GPU while another
    GPU idles: measured 55 % busy GPUs and 106 GB on one device at BASELINE configs[4]).  The host's cores are shared out between
    the workers: with every worker at torch's default (all cores) eight of them spent more time contending for 16 threads
    than fitting -- 2.2 s per config-5 fit against 0.9 s (the fits are GPU-bound; the host side is small tensor bookkeeping)."""
    global _WORKER_DEVICE
    _WORKER_DEVICE = device_queue.get()
    try:
        import torch
        torch.set_num_threads(max(1, (multiprocessing.cpu_count() or 1) // max(1, int(n_workers))))
    except Exception:       # pragma: no cover - thread pools already fixed by the embedding application
        pass


def _call_on_worker_device(task_fn, device_arg, args):
    args = list(args)
    if _WORKER_DEVICE is not None and device_arg is not None:
        args[device_arg] = _WORKER_DEVICE
    return task_fn(*args)


def execute_task_in_parallel(task_fn, inputs, max_workers=None, devices=None, device_arg=None):
    """``inputs``: {key: args tuple}; returns {key: task_fn(*args)}.

    ``devices`` + ``device_arg`` (extension): one worker per device, each bound to its device for good; the task's argument at
    position ``device_arg`` is replaced by the worker's device, so jobs are pulled by whichever GPU becomes free."""
    max_workers = max_workers or multiprocessing.cpu_count()
    if devices is not None:
        max_workers = min(max_workers, len(devices))
    if max_workers <= 1 or len(inputs) <= 1:
        return {key: task_fn(*args) for key, args in inputs.items()}
    ctx = multiprocessing.get_context("spawn")
    results = {}
    if devices is not None and device_arg is not None:
        q = ctx.Queue()
        for d in list(devices)[:max_workers]:
            q.put(d)
        pool_kw = dict(initializer=_bind_worker, initargs=(q, max_workers))
        submit = lambda pool, args: pool.submit(_call_on_worker_device, task_fn, device_arg, args)  # noqa: E731
    else:
        pool_kw = {}
        submit = lambda pool, args: pool.submit(task_fn, *args)  # noqa: E731
    with concurrent.futures.ProcessPoolExecutor(max_workers, mp_context=ctx, **pool_kw) as pool:
        futures = {submit(pool, args): key for key, args in inputs.items()}
        for fut in concurrent.futures.as_completed(futures):
            results[futures[fut]] = fut.result()
    return results
