"""JobHandler classes of the reference (treeCl/parutils.py:213-283), kept so that call sites compile unchanged.

In the reference they are the only parallelism of the distance path: a task is mapped over N(N-1)/2 argument tuples
sequentially, on a thread pool or on a process pool (parutils.py:113-210).  Here the whole matrix is one GPU job, so
`jobhandler` and `batchsize` are accepted by Collection.get_inter_tree_distances and ignored.  The handlers still work
as plain mappers for callers that use them directly.
"""
import concurrent.futures


class JobHandler(object):
    def __call__(self, task, args, message, batchsize, nargs=None):
        raise NotImplementedError


class SequentialJobHandler(JobHandler):
    def __call__(self, task, args, message, batchsize, nargs=None):
        return [task(*a) for a in args]


class ThreadpoolJobHandler(JobHandler):
    def __init__(self, concurrency):
        self.concurrency = concurrency

    def __call__(self, task, args, message, batchsize, nargs=None):
        with concurrent.futures.ThreadPoolExecutor(max_workers=self.concurrency) as ex:
            return list(ex.map(lambda a: task(*a), args))


class ProcesspoolJobHandler(ThreadpoolJobHandler):
    """The GPU engine is not fork-safe; a process pool adds nothing here, so this maps on threads."""


default_jobhandler = SequentialJobHandler()
