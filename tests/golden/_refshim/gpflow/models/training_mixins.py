"""gpflow.models.training_mixins.InternalDataTrainingLossMixin: training_loss = -(MLE objective + log prior); no priors here."""


class InternalDataTrainingLossMixin:
    def training_loss(self):
        return self._training_loss()

    def training_loss_closure(self, *, compile=True):
        return self.training_loss
