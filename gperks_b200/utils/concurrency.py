"""Task-level parallelism for independent fits (role of GPErks/utils/concurrency.py): a process pool with the
'spawn' start method (CUDA contexts do not survive fork), or in-process execution when one worker suffices."""
import concurrent.futures
import multiprocessing


def execute_task_in_parallel(task_fn, inputs, max_workers=None):
    """``inputs``: {key: args tuple}; returns {key: task_fn(*args)}."""
    max_workers = max_workers or multiprocessing.cpu_count()
    if max_workers <= 1 or len(inputs) <= 1:
        return {key: task_fn(*args) for key, args in inputs.items()}
    ctx = multiprocessing.get_context("spawn")
    results = {}
    with concurrent.futures.ProcessPoolExecutor(max_workers, mp_context=ctx) as pool:
        futures = {pool.submit(task_fn, *args): key for key, args in inputs.items()}
        for fut in concurrent.futures.as_completed(futures):
            results[futures[fut]] = fut.result()
    return results
