class DictConfig(dict):
    pass


class OmegaConf:
    @staticmethod
    def structured(x):
        return x

    @staticmethod
    def create(x):
        return x
