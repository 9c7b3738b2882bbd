"""Base of Interaction (data/data.py): the run's configuration plus the raw training / test triples."""


class Data(object):
    def __init__(self, conf, training, test):
        self.config, self.training_data, self.test_data = conf, training, test
