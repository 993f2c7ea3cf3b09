"""Logging helper with the extra VERBOSE level the reference's model uses (reference `_log.py:10-46`)."""
import logging

VERBOSE = 5
logging.addLevelName(VERBOSE, "VERBOSE")


def _verbose(self, message, *args, **kws):
    if self.isEnabledFor(VERBOSE):
        self._log(VERBOSE, message, args, **kws)


logging.Logger.verbose = _verbose


def setup_logging(name, verbose=False):
    logger = logging.getLogger(name)
    logger.setLevel(logging.DEBUG if verbose else logging.INFO)
    if not logger.handlers:
        h = logging.StreamHandler()
        h.setFormatter(logging.Formatter("%(asctime)s - [PID %(process)d] - %(name)-25s - %(levelname)s - %(message)s"))
        logger.addHandler(h)
    return logger
