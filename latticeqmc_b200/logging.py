"""Host-side text logging (console + ``lqmc.log``), API-compatible with the reference
``lqmc/logging.py:1-72``.  Never on a timed path: the engine's per-proposal record is the
optional device trace buffer (``lqmc_get_trace``), not a log line per proposal."""
import logging
from logging import getLogger, INFO, DEBUG, WARNING, ERROR  # noqa: F401  (re-exported levels)

FILE = "lqmc.log"
BLACK, RED, GREEN, YELLOW, BLUE, MAGENTA, CYAN, WHITE = range(8)


class ConsoleFormatter(logging.Formatter):

    COLORS = {"DEBUG": WHITE, "INFO": WHITE, "WARNING": YELLOW, "ERROR": RED, "CRITICAL": RED}

    def __init__(self, fmt, color=True, datefmt=None):
        super().__init__(fmt, datefmt)
        self.color = color

    def format(self, record):
        text = super().format(record)
        if not self.color:
            return text
        code = 30 + self.COLORS.get(record.levelname, WHITE)
        bold = "\033[1m" if record.levelname == "CRITICAL" else ""
        return f"\033[1;{code}m{bold}{text}\033[0m"


class ConsoleHandler(logging.StreamHandler):

    def __init__(self, level=logging.INFO, lvlname=True):
        super().__init__()
        self.setFormatter(ConsoleFormatter("[%(levelname)-5s] %(message)s" if lvlname else "%(message)s"))
        self.setLevel(level)


class FileHandler(logging.FileHandler):

    def __init__(self, filename, level=logging.DEBUG, mode="w", datefmt="%H:%M:%S"):
        super().__init__(filename, mode=mode)
        self.setFormatter(logging.Formatter("[%(levelname)-5s] %(asctime)s - %(message)s", datefmt))
        self.setLevel(level)


def get_logger(name="lqmc", file=FILE, console_lvl=INFO, file_lvl=DEBUG):
    logger = getLogger(name)
    if not logger.handlers:     # the reference adds a fresh pair on every call; once is enough
        logger.addHandler(ConsoleHandler(console_lvl))
        logger.addHandler(FileHandler(file, file_lvl))
    return logger


def read_log_file(file=FILE):
    with open(file, "r") as fh:
        return [line.rstrip("\n") for line in fh]
