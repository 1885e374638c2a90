"""Stand-alone `experiment` object for running the models without the reference's `mag` experiment manager.

The reference's model classes take an `experiment` (networks/classifiers.py:485-491): they read nested attributes of
`experiment.config.{data,network,train}`, call `experiment.register_directory(name)` and use
`experiment.{checkpoints,summaries}`.  `StandaloneExperiment` provides exactly that over a plain nested dict, and
`make_config` builds the dict with the keys those classes read (defaults = the README's documented
training command, README.md:101-131)."""
import os


class AttrDict(dict):
    """Nested dict with attribute access (what `mag` gives the reference)."""

    def __getattr__(self, key):
        try:
            value = self[key]
        except KeyError:
            raise AttributeError(key)
        return AttrDict(value) if isinstance(value, dict) else value


class StandaloneExperiment:
    def __init__(self, config, root="/tmp/fsb200_experiment"):
        self.config = AttrDict(config)
        self.root = root
        self.checkpoints = os.path.join(root, "checkpoints")
        self.predictions = os.path.join(root, "predictions")

    def register_directory(self, name):
        path = os.path.join(self.root, name)
        os.makedirs(path, exist_ok=True)
        setattr(self, name, path)

    def register_result(self, *args, **kwargs):
        pass


def make_config(features="mel_2048_1024_128", num_conv_blocks=5, conv_base_depth=100, growth_rate=1.5,
                start_deep_supervision_on=1, output_dropout=0.0, n_classes=80, input_dim=None, aggregation_type="max",
                scheduler="1cycle_0.0001_0.005", weight_decay=0.0, accumulation_steps=1, learning_rate=0.001,
                optimizer="adam"):
    if input_dim is None:
        kind, *args = features.split("_")
        input_dim = int(args[2]) if kind == "mel" else int(args[0]) // 2 + 1
    return dict(
        data=dict(features=features, _input_dim=input_dim, _n_classes=n_classes),
        network=dict(num_conv_blocks=num_conv_blocks, start_deep_supervision_on=start_deep_supervision_on,
                     conv_base_depth=conv_base_depth, growth_rate=growth_rate, output_dropout=output_dropout,
                     aggregation_type=aggregation_type),
        train=dict(accumulation_steps=accumulation_steps, learning_rate=learning_rate, optimizer=optimizer,
                   scheduler=scheduler, weight_decay=weight_decay, _save_every=1000,
                   switch_off_augmentations_on=1000))
