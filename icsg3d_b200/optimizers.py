"""keras.optimizers.Adam stand-in: a hyper-parameter holder; the update runs in icsg3d_adam_keras_step."""


class Adam:
    def __init__(self, lr=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        self.lr, self.beta_1, self.beta_2, self.epsilon = float(lr), float(beta_1), float(beta_2), float(epsilon)
