"""oracle/mnet_oracle.py — TEST INFRASTRUCTURE (never imported by the product): numpy restatement of the reference's recurrent
stack, sample by sample in the reference's order.

  RMLayer::act            NeuralNet/RMLayer.inl:165-187     h_t = f(W_in x_t + b + W_rec h_{t-1})
  RMLayer::backpropagate  NeuralNet/RMLayer.inl:189-233     a = add_in_grad(dC/dh_t, f'_t, a); db += a; dW_in += a x_t^T;
                                                            dx_t = W_in^T a; dW_rec += a h_{t-1}^T; a = (W_rec^T a) * f'_{t-1}
  RMLayer::update         NeuralNet/RMLayer.inl:235-260     gradient-norm clipping of the learning rate, biases not regularised
  MNet::add_gradient      NeuralNet/MNet.inl:123-146        forward with traces, cost derivative at every time step, backward
  MNet::learn             NeuralNet/MNet.inl:171-180        gradient SUM over the items, update with -learning_rate / n
  activations             Math/ActivationFunctionHelper.cpp:64-114,  cost derivatives  Math/CostFunctionFactory.h:29-55

Pinned against the reference itself: tests/test_mnet_oracle.py runs it beside oracle/_ref/libdlrefm_{f32,f64}.so (the unmodified
MNet<CpuDC> / RMLayer<CpuDC>, oracle/ref_mnet_driver.cpp) and against the reference-generated fixtures tests/golden/mnet_*.npz.
Flat parameter layout per layer: bias [H] | W_in [H][I] | W_rec [H][H]."""
import numpy as np

SIGMOID, TANH, RELU, SOFTMAX, LINEAR = 1, 2, 3, 4, 5
SQUARED_ERROR, CROSS_ENTROPY = 1, 2


def _activate(act, z):
    if act == SIGMOID:
        v = (1 / (1 + np.exp(-z))).astype(z.dtype); return v, v * (1 - v)
    if act == TANH:
        v = np.tanh(z); return v, 1 - v * v
    if act == RELU:
        return np.where(z > 0, z, 0).astype(z.dtype), (z > 0).astype(z.dtype)
    if act == LINEAR:
        return z.copy(), np.ones_like(z)
    raise ValueError("element-wise activations only")


def _cost_derivative(cost_id, x, ref):
    if cost_id == SQUARED_ERROR:
        return x - ref
    if cost_id == CROSS_ENTROPY:
        with np.errstate(divide="ignore", invalid="ignore"):
            d = np.where(ref <= 0, 1 / (1 - x), np.where(ref >= 1, -1 / x, -(ref / x - (1 - ref) / (1 - x)))).astype(x.dtype)
            val = np.where(ref <= 0, -np.log(1 - x), np.where(ref >= 1, -np.log(x), -(ref * np.log(x) + (1 - ref) * np.log(1 - x))))
        return np.where(np.isfinite(d) & np.isfinite(val), d, 0).astype(x.dtype)
    raise ValueError("unknown cost function")


class MNetOracle:
    def __init__(self, depth, dims, activations, clips, real="f32"):
        self.depth, self.dims, self.acts, self.clips = int(depth), list(dims), list(activations), list(clips)
        self.dtype = np.float32 if real == "f32" else np.float64
        self.offsets = []
        off = 0
        for l in range(len(self.acts)):
            I, H = self.dims[l], self.dims[l + 1]
            self.offsets.append(off); off += H + H * I + H * H
        self.param_count = off
        self.params = np.zeros(off, self.dtype)

    def _views(self, flat, l):
        I, H = self.dims[l], self.dims[l + 1]
        o = self.offsets[l]
        return flat[o:o + H], flat[o + H:o + H + H * I].reshape(H, I), flat[o + H + H * I:o + H + H * I + H * H].reshape(H, H)

    def set_params(self, p): self.params = np.ascontiguousarray(p, self.dtype).copy()
    def get_params(self): return self.params.copy()

    def _forward(self, x):   # x [T][I0] -> per layer (inputs, outputs, derivatives)
        trace = []
        cur = x
        for l, act in enumerate(self.acts):
            b, Win, Wrec = self._views(self.params, l)
            H = np.zeros((self.depth, self.dims[l + 1]), self.dtype); D = np.zeros_like(H)
            for t in range(self.depth):
                z = Win @ cur[t] + b
                if t > 0: z = Wrec @ H[t - 1] + z
                H[t], D[t] = _activate(act, z.astype(self.dtype))
            trace.append((cur, H, D)); cur = H
        return trace

    def act(self, xs):
        xs = np.ascontiguousarray(xs, self.dtype)
        return np.stack([self._forward(x)[-1][1] for x in xs])

    def calc_gradient_sum(self, xs, ys, cost_id):
        xs = np.ascontiguousarray(xs, self.dtype); ys = np.ascontiguousarray(ys, self.dtype)
        g = np.zeros(self.param_count, self.dtype)
        for x, y in zip(xs, ys):
            trace = self._forward(x)
            delta = _cost_derivative(cost_id, trace[-1][1], y)           # [T][out]
            for l in range(len(self.acts) - 1, -1, -1):
                inp, H, D = trace[l]
                b, Win, Wrec = self._views(self.params, l)
                gb, gWin, gWrec = self._views(g, l)
                a = np.zeros(self.dims[l + 1], self.dtype)
                dx = np.zeros_like(inp)
                for t in range(self.depth - 1, -1, -1):
                    a = a + delta[t] * D[t]
                    gb += a; gWin += np.outer(a, inp[t])
                    if l > 0: dx[t] = Win.T @ a
                    if t > 0:
                        gWrec += np.outer(a, H[t - 1])
                        a = ((Wrec.T @ a) * D[t - 1]).astype(self.dtype)
                delta = dx
        return g

    def learn(self, xs, ys, cost_id, learning_rate, reg_factor=0.0):
        g = self.calc_gradient_sum(xs, ys, cost_id)
        R = self.dtype
        scale = R(-R(learning_rate) / R(len(xs)))
        for l in range(len(self.acts)):
            I, H = self.dims[l], self.dims[l + 1]
            o = self.offsets[l]; n = H + H * I + H * H
            eff = scale
            clip = R(self.clips[l])
            if clip > 0:
                norm = R(np.sqrt(np.sum(g[o:o + n].astype(np.float64) ** 2)))
                if norm > clip: eff = R(eff * (clip / norm))
            self.params[o:o + H] += g[o:o + H] * eff
            if reg_factor <= 0: self.params[o + H:o + n] += g[o + H:o + n] * eff
            else: self.params[o + H:o + n] = self.params[o + H:o + n] * R(R(1) + R(reg_factor) * eff) + g[o + H:o + n] * eff
