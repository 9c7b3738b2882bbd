class Data(object):
    """Mirror of data/data.py: holds the config and the raw train/test triples."""

    def __init__(self, conf, training, test):
        self.config = conf
        self.training_data = training
        self.test_data = test
