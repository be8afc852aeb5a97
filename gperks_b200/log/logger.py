"""Process-wide logger: INFO to the console, DEBUG to ~/gperks_b200.log when writable
(GPErks/log/logger.py:10-34)."""
import logging

from ..constants import APPLICATION_NAME, DEFAULT_LOG_FILE, DEFAULT_LOG_FORMAT


def get_logger(name=APPLICATION_NAME, log_format=DEFAULT_LOG_FORMAT, stdout_level=logging.INFO,
               file_name=DEFAULT_LOG_FILE, file_level=logging.DEBUG):
    logger = logging.getLogger(name)
    if logger.handlers:
        return logger
    logger.setLevel(logging.DEBUG)
    fmt = logging.Formatter(log_format)
    console = logging.StreamHandler()
    console.setLevel(stdout_level)
    console.setFormatter(fmt)
    logger.addHandler(console)
    try:
        fh = logging.FileHandler(file_name)
        fh.setLevel(file_level)
        fh.setFormatter(fmt)
        logger.addHandler(fh)
    except OSError:
        pass
    return logger
