"""Millisecond timestamps and their printable form (GPErks/utils/time.py:1-23)."""
import time as _time
from datetime import datetime

DEFAULT_DATE_FORMAT = "%d-%m-%Y_%Hh%Mm%Ss"

SECOND = 1000
MINUTE, HOUR = 60 * SECOND, 3600 * SECOND
DAY = 24 * HOUR
WEEK, MONTH, YEAR = 7 * DAY, 30 * DAY, 365 * DAY


def now() -> int:
    """Milliseconds since the epoch."""
    return int(round(_time.time() * 1000))


def pretty_str(now_ms: int, date_format=DEFAULT_DATE_FORMAT) -> str:
    return datetime.fromtimestamp(now_ms / 1000.0).strftime(date_format)
