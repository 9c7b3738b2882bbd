"""Mirror of util/logger.py:6-17 (Log): ./log/<filename>.log relative to the cwd."""
import logging
import os


class Log(object):
    def __init__(self, module, filename):
        self.logger = logging.getLogger(module)
        self.logger.setLevel(level=logging.INFO)
        os.makedirs("./log/", exist_ok=True)
        handler = logging.FileHandler("./log/" + filename + ".log")
        handler.setFormatter(logging.Formatter("%(asctime)s - %(name)s - %(levelname)s - %(message)s"))
        self.logger.addHandler(handler)

    def add(self, text):
        self.logger.info(text)
